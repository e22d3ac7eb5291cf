/*
 * mrts_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY; see mrts_oracle.h).
 *
 * Restates, in plain C, the Java game rules of ConnAALL/MicroRTS.  Every function names the
 * reference file:line it follows.  Data structures deliberately mirror the Java ones (ordered
 * unit list with linear scans, insertion-ordered assignment map, heap "objects" that stay
 * addressable after removal) so that the restatement can be audited line by line; this is the
 * opposite of how the CUDA engine is organised, which keeps the two implementations independent.
 */
#include "mrts_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------
 * java.util.Random (Java SE specification; JDK class, not in /root/reference -- see SURVEY 8c)
 * ---------------------------------------------------------------------------------------------- */
#define JR_MULT 0x5DEECE66DULL
#define JR_ADD 0xBULL
#define JR_MASK ((1ULL << 48) - 1)

void o_jr_seed(OJRandom *r, int64_t seed) { r->s = ((uint64_t)seed ^ JR_MULT) & JR_MASK; }
int32_t o_jr_next(OJRandom *r, int bits) {
    r->s = (r->s * JR_MULT + JR_ADD) & JR_MASK;
    return (int32_t)(int64_t)(r->s >> (48 - bits)); /* (int)(seed >>> (48-bits)) */
}
int32_t o_jr_next_int(OJRandom *r) { return o_jr_next(r, 32); }
int32_t o_jr_next_int_bound(OJRandom *r, int32_t bound) {
    int32_t v = o_jr_next(r, 31);
    int32_t m = bound - 1;
    if ((bound & m) == 0) {
        v = (int32_t)(((int64_t)bound * (int64_t)v) >> 31);
    } else {
        int32_t u = v;
        /* for (int u = r; u - (r = u % bound) + m < 0; u = next(31)); -- with Java int overflow */
        for (;;) {
            v = u % bound;
            int32_t t = (int32_t)((uint32_t)u - (uint32_t)v + (uint32_t)m);
            if (t >= 0) break;
            u = o_jr_next(r, 31);
        }
    }
    return v;
}
double o_jr_next_double(OJRandom *r) {
    int64_t a = (int64_t)o_jr_next(r, 26);
    int64_t b = (int64_t)o_jr_next(r, 27);
    return (double)((a << 27) + b) * 0x1.0p-53;
}

/* ------------------------------------------------------------------------------------------------
 * UnitTypeTable  (src/rts/units/UnitTypeTable.java:104-289, UnitType.java:23-110)
 * ---------------------------------------------------------------------------------------------- */
#define MAXT 16
struct OUtt {
    int n;
    int conflict;
    int16_t f[MAXT][OF_NFIELDS];
    int flags[MAXT];
    int nprod[MAXT];
    uint8_t prod[MAXT][MAXT];
};

static void utt_defaults(OUtt *t, int id) {
    /* UnitType.java:23-110 field initialisers */
    int16_t *f = t->f[id];
    f[OF_COST] = 1; f[OF_HP] = 1; f[OF_MINDMG] = 1; f[OF_MAXDMG] = 1; f[OF_RANGE] = 1;
    f[OF_PRODUCE_T] = 10; f[OF_MOVE_T] = 10; f[OF_ATTACK_T] = 10; f[OF_HARVEST_T] = 10; f[OF_RETURN_T] = 10;
    f[OF_HARVEST_AMT] = 1; f[OF_SIGHT] = 4;
    t->flags[id] = OFL_MOVE | OFL_ATTACK;
    t->nprod[id] = 0;
}

OUtt *o_utt_empty(int conflict) {
    OUtt *t = (OUtt *)calloc(1, sizeof(OUtt));
    t->conflict = conflict;
    return t;
}

OUtt *o_utt_create(int version, int conflict) {
    /* UnitTypeTable.setUnitTypeTable, UnitTypeTable.java:104-289; version 1 ORIGINAL, 2 FINETUNED, 3 NON_DETERMINISTIC */
    OUtt *t = o_utt_empty(conflict);
    t->n = 7;
    for (int i = 0; i < 7; i++) utt_defaults(t, i);
    int16_t *f;
    /* Resource :108-117 */
    f = t->f[0]; t->flags[0] = OFL_RESOURCE; f[OF_SIGHT] = 0;
    /* Base :120-138 (v3 leaves produceTime at the default 10) */
    f = t->f[1]; f[OF_COST] = 10; f[OF_HP] = 10;
    if (version == 1) f[OF_PRODUCE_T] = 250; else if (version == 2) f[OF_PRODUCE_T] = 200;
    t->flags[1] = OFL_STOCKPILE; f[OF_SIGHT] = 5;
    /* Barracks :141-161 */
    f = t->f[2]; f[OF_COST] = 5; f[OF_HP] = 4;
    if (version == 1) f[OF_PRODUCE_T] = 200; else if (version == 2 || version == 3) f[OF_PRODUCE_T] = 100;
    t->flags[2] = 0; f[OF_SIGHT] = 3;
    /* Worker :164-190 */
    f = t->f[3]; f[OF_COST] = 1; f[OF_HP] = 1;
    if (version == 3) { f[OF_MINDMG] = 0; f[OF_MAXDMG] = 2; } else { f[OF_MINDMG] = f[OF_MAXDMG] = 1; }
    f[OF_RANGE] = 1; f[OF_PRODUCE_T] = 50; f[OF_MOVE_T] = 10; f[OF_ATTACK_T] = 5; f[OF_HARVEST_T] = 20; f[OF_RETURN_T] = 10;
    t->flags[3] = OFL_HARVEST | OFL_MOVE | OFL_ATTACK; f[OF_SIGHT] = 3;
    /* Light :193-216 */
    f = t->f[4]; f[OF_COST] = 2; f[OF_HP] = 4;
    if (version == 3) { f[OF_MINDMG] = 1; f[OF_MAXDMG] = 3; } else { f[OF_MINDMG] = f[OF_MAXDMG] = 2; }
    f[OF_RANGE] = 1; f[OF_PRODUCE_T] = 80; f[OF_MOVE_T] = 8; f[OF_ATTACK_T] = 5;
    t->flags[4] = OFL_MOVE | OFL_ATTACK; f[OF_SIGHT] = 2;
    /* Heavy :219-254 */
    f = t->f[5];
    if (version == 3) { f[OF_MINDMG] = 0; f[OF_MAXDMG] = 6; } else { f[OF_MINDMG] = f[OF_MAXDMG] = 4; }
    f[OF_RANGE] = 1; f[OF_PRODUCE_T] = 120;
    if (version == 1) { f[OF_MOVE_T] = 12; f[OF_HP] = 4; f[OF_COST] = 2; }
    else if (version == 2 || version == 3) { f[OF_MOVE_T] = 10; f[OF_HP] = 8; f[OF_COST] = 3; }
    f[OF_ATTACK_T] = 5; t->flags[5] = OFL_MOVE | OFL_ATTACK; f[OF_SIGHT] = 2;
    /* Ranged :257-279 */
    f = t->f[6]; f[OF_COST] = 2; f[OF_HP] = 1;
    if (version == 3) { f[OF_MINDMG] = 1; f[OF_MAXDMG] = 2; } else { f[OF_MINDMG] = f[OF_MAXDMG] = 1; }
    f[OF_RANGE] = 3; f[OF_PRODUCE_T] = 100; f[OF_MOVE_T] = 10; f[OF_ATTACK_T] = 5;
    t->flags[6] = OFL_MOVE | OFL_ATTACK; f[OF_SIGHT] = 3;
    /* produces :282-288 */
    t->nprod[1] = 1; t->prod[1][0] = 3;
    t->nprod[2] = 3; t->prod[2][0] = 4; t->prod[2][1] = 5; t->prod[2][2] = 6;
    t->nprod[3] = 2; t->prod[3][0] = 1; t->prod[3][1] = 2;
    return t;
}

void o_utt_set_type(OUtt *t, int id, const int16_t fields[OF_NFIELDS], int flags, int nprod, const uint8_t *prod) {
    if (id >= t->n) t->n = id + 1;
    memcpy(t->f[id], fields, sizeof(int16_t) * OF_NFIELDS);
    t->flags[id] = flags;
    t->nprod[id] = nprod;
    for (int i = 0; i < nprod; i++) t->prod[id][i] = prod[i];
}
int o_utt_field(const OUtt *t, int id, int field) { return t->f[id][field]; }
int o_utt_flags(const OUtt *t, int id) { return t->flags[id]; }
int o_utt_produces(const OUtt *t, int id, uint8_t *out) {
    for (int i = 0; i < t->nprod[id]; i++) out[i] = t->prod[id][i];
    return t->nprod[id];
}
int o_utt_max_attack_range(const OUtt *t) {
    /* UnitTypeTable.getMaxAttackRange */
    int m = 0;
    for (int i = 0; i < t->n; i++) if (t->f[i][OF_RANGE] > m) m = t->f[i][OF_RANGE];
    return m;
}
void o_utt_free(OUtt *t) { free(t); }

/* ------------------------------------------------------------------------------------------------
 * State.  Units are "heap objects" in a pool (index = identity, never reused), the unit list holds pool
 * indices in list order (PhysicalGameState.units, a LinkedList), assignments are kept in insertion
 * order (GameState.unitActions, a LinkedHashMap).                      GameState.java:34-46
 * ---------------------------------------------------------------------------------------------- */
typedef struct { int type, player, x, y, res, hp; int64_t id; } OUnit;

typedef struct { int npos; int pos; int res[2]; } ORu1; /* ResourceUsage of ONE action: at most one position */

typedef struct {
    int type, param, x, y, utype;
    ORu1 ru; /* r_cache (UnitAction.java:130,246-296), filled by act_ru() */
    int ru_done;
} OAct;

typedef struct { int unit; OAct act; int time; } OAssign;

/* merged ResourceUsage (ResourceUsage.java:12-13): list of positions + per-player resources */
typedef struct { int npos; int cap; int *pos; int res[2]; } ORu;

struct OGame {
    const OUtt *utt;
    int w, h;
    uint8_t *terrain;
    int res[2];
    OUnit *pool; int pool_n, pool_cap;
    int *list; int n, list_cap;
    OAssign *asg; int na, asg_cap;
    int time;
    int cancel_ctr;
    int64_t next_id;
    OJRandom rng_policy;   /* util.Sampler.generator   (Sampler.java:17)  */
    OJRandom rng_conflict; /* GameState.r               (GameState.java:37) */
    OJRandom rng_damage;   /* UnitAction.r              (UnitAction.java:24) */
    int errors;
    int po_observer;       /* 1 + the observer when this is a PartiallyObservableGameState view (o_po_view), else 0 */
};

enum { OE_ADD_OCCUPIED = 1, OE_MIXED_OWNERS = 2, OE_BAD_UNIT = 4, OE_INCONSISTENT_OLDER = 8, OE_FAILED_PRODUCE = 16,
       OE_BAD_ACTION = 32 };

static void ru_init(ORu *r) { r->npos = 0; r->cap = 0; r->pos = NULL; r->res[0] = r->res[1] = 0; }
static void ru_free(ORu *r) { free(r->pos); r->pos = NULL; r->npos = r->cap = 0; }
static void ru_add_pos(ORu *r, int p) {
    if (r->npos == r->cap) { r->cap = r->cap ? r->cap * 2 : 16; r->pos = (int *)realloc(r->pos, sizeof(int) * r->cap); }
    r->pos[r->npos++] = p;
}
static int ru_contains(const ORu *r, int p) {
    for (int i = 0; i < r->npos; i++) if (r->pos[i] == p) return 1;
    return 0;
}
/* ResourceUsage.merge, ResourceUsage.java:93-98 */
static void ru_merge1(ORu *r, const ORu1 *o) {
    if (o->npos) ru_add_pos(r, o->pos);
    r->res[0] += o->res[0]; r->res[1] += o->res[1];
}

/* ResourceUsage.consistentWith(anotherUsage), ResourceUsage.java:31-50.  Four flavours by operand shape. */
static int res_consistent(const int a[2], const int b[2], const OGame *g) {
    for (int i = 0; i < 2; i++) {
        if (b[i] == 0) continue;
        if (a[i] + b[i] > 0 && a[i] + b[i] > g->res[i]) return 0;
    }
    return 1;
}
static int ru1_consistent_with_ru1(const ORu1 *self, const ORu1 *other, const OGame *g) {
    if (other->npos && self->npos && self->pos == other->pos) return 0;
    return res_consistent(self->res, other->res, g);
}
static int ru1_consistent_with_ru(const ORu1 *self, const ORu *other, const OGame *g) {
    for (int i = 0; i < other->npos; i++) if (self->npos && self->pos == other->pos[i]) return 0;
    return res_consistent(self->res, other->res, g);
}
static int ru_consistent_with_ru1(const ORu *self, const ORu1 *other, const OGame *g) {
    if (other->npos && ru_contains(self, other->pos)) return 0;
    return res_consistent(self->res, other->res, g);
}

static const int DX[4] = {0, 1, 0, -1}, DY[4] = {-1, 0, 1, 0}; /* UnitAction.java:94,100 */

OGame *o_game_create(const OUtt *utt, int w, int h, const uint8_t *terrain, int res0, int res1) {
    OGame *g = (OGame *)calloc(1, sizeof(OGame));
    g->utt = utt; g->w = w; g->h = h;
    g->terrain = (uint8_t *)malloc((size_t)w * h);
    memcpy(g->terrain, terrain, (size_t)w * h);
    g->res[0] = res0; g->res[1] = res1;
    o_game_seed(g, 0);
    return g;
}

static int pool_new(OGame *g) {
    if (g->pool_n == g->pool_cap) { g->pool_cap = g->pool_cap ? g->pool_cap * 2 : 64; g->pool = (OUnit *)realloc(g->pool, sizeof(OUnit) * g->pool_cap); }
    return g->pool_n++;
}
static void list_append(OGame *g, int u) {
    if (g->n == g->list_cap) { g->list_cap = g->list_cap ? g->list_cap * 2 : 64; g->list = (int *)realloc(g->list, sizeof(int) * g->list_cap); }
    g->list[g->n++] = u;
}

/* PhysicalGameState.getUnitAt, PhysicalGameState.java:263-270 (first match in list order) */
static int unit_at(const OGame *g, int x, int y) {
    for (int i = 0; i < g->n; i++) { const OUnit *u = &g->pool[g->list[i]]; if (u->x == x && u->y == y) return g->list[i]; }
    return -1;
}
static int list_index_of(const OGame *g, int u) {
    for (int i = 0; i < g->n; i++) if (g->list[i] == u) return i;
    return -1;
}

/* PhysicalGameState.addUnit, PhysicalGameState.java:189-201 */
static int add_unit(OGame *g, int u) {
    const OUnit *nu = &g->pool[u];
    for (int i = 0; i < g->n; i++) {
        const OUnit *e = &g->pool[g->list[i]];
        if (e->x == nu->x && e->y == nu->y) { g->errors |= OE_ADD_OCCUPIED; return 0; } /* Java throws */
    }
    list_append(g, u);
    return 1;
}

void o_game_add_unit(OGame *g, int type, int64_t id, int player, int x, int y, int res, int hp) {
    /* Unit(long ID,...) ctor, Unit.java:72-83 (hitpoints overridden by the XML value, Unit.fromXML) */
    int u = pool_new(g);
    OUnit *p = &g->pool[u];
    p->type = type; p->player = player; p->x = x; p->y = y; p->res = res; p->hp = hp; p->id = id;
    if (id >= g->next_id) g->next_id = id + 1;
    add_unit(g, u);
}

OGame *o_game_clone(const OGame *s) {
    /* GameState.clone, GameState.java:591-610: deep copy; assignment order preserved */
    OGame *g = (OGame *)malloc(sizeof(OGame));
    *g = *s;
    g->terrain = (uint8_t *)malloc((size_t)s->w * s->h); memcpy(g->terrain, s->terrain, (size_t)s->w * s->h);
    g->pool_cap = s->pool_n + 64; g->pool = (OUnit *)malloc(sizeof(OUnit) * g->pool_cap); memcpy(g->pool, s->pool, sizeof(OUnit) * s->pool_n);
    g->list_cap = s->n + 64; g->list = (int *)malloc(sizeof(int) * g->list_cap); memcpy(g->list, s->list, sizeof(int) * s->n);
    g->asg_cap = s->na + 64; g->asg = (OAssign *)malloc(sizeof(OAssign) * g->asg_cap); memcpy(g->asg, s->asg, sizeof(OAssign) * s->na);
    return g;
}
void o_game_free(OGame *g) { if (!g) return; free(g->terrain); free(g->pool); free(g->list); free(g->asg); free(g); }
void o_game_seed(OGame *g, int64_t seed) {
    /* One stream per static Random of the reference.  The reference seeds none of them (parity unpinned);
     * our convention: policy stream = Random(seed), conflict = Random(seed ^ 0x5851F42D4C957F2D),
     * damage = Random(seed ^ 0x14057B7EF767814F). */
    o_jr_seed(&g->rng_policy, seed);
    o_jr_seed(&g->rng_conflict, seed ^ 0x5851F42D4C957F2DLL);
    o_jr_seed(&g->rng_damage, seed ^ 0x14057B7EF767814FLL);
}
int o_game_time(const OGame *g) { return g->time; }
int o_game_n_units(const OGame *g) { return g->n; }
int o_game_resources(const OGame *g, int p) { return g->res[p]; }
int o_game_errors(const OGame *g) { return g->errors; }
void o_game_set_rng_state(OGame *g, int which, int64_t s) { OJRandom *r = which == 0 ? &g->rng_policy : (which == 1 ? &g->rng_conflict : &g->rng_damage); r->s = (uint64_t)s; }
int64_t o_game_rng_state(const OGame *g, int which) { return (int64_t)(which == 0 ? g->rng_policy.s : (which == 1 ? g->rng_conflict.s : g->rng_damage.s)); }

int o_game_units(const OGame *g, int32_t *out) {
    for (int i = 0; i < g->n; i++) {
        const OUnit *u = &g->pool[g->list[i]];
        int32_t *o = out + i * 8;
        o[0] = u->type; o[1] = u->player; o[2] = u->x; o[3] = u->y; o[4] = u->res; o[5] = u->hp;
        o[6] = (int32_t)(u->id & 0xffffffff); o[7] = (int32_t)(u->id >> 32);
    }
    return g->n;
}

static int find_assign(const OGame *g, int u) {
    for (int i = 0; i < g->na; i++) if (g->asg[i].unit == u) return i;
    return -1;
}

int o_game_assignments(const OGame *g, int32_t *out) {
    for (int i = 0; i < g->n; i++) {
        int32_t *o = out + i * 8;
        int a = find_assign(g, g->list[i]);
        memset(o, 0, sizeof(int32_t) * 8);
        if (a >= 0) {
            const OAssign *s = &g->asg[a];
            o[0] = 1; o[1] = s->act.type; o[2] = s->act.param; o[3] = s->act.x; o[4] = s->act.y; o[5] = s->act.utype;
            o[6] = s->time; o[7] = a;
        }
    }
    return g->n;
}

/* PhysicalGameState.winner / gameover, PhysicalGameState.java:334-387 */
int o_game_winner(const OGame *g) {
    int cnt[2] = {0, 0};
    for (int i = 0; i < g->n; i++) { int p = g->pool[g->list[i]].player; if (p >= 0) cnt[p]++; }
    int winner = -1;
    for (int i = 0; i < 2; i++) if (cnt[i] > 0) { if (winner == -1) winner = i; else return -1; }
    return winner;
}
int o_game_gameover(const OGame *g) {
    int cnt[2] = {0, 0}, total = 0;
    for (int i = 0; i < g->n; i++) { int p = g->pool[g->list[i]].player; if (p >= 0) { cnt[p]++; total++; } }
    if (total == 0) return 1;
    int winner = -1;
    for (int i = 0; i < 2; i++) if (cnt[i] > 0) { if (winner == -1) winner = i; else return 0; }
    return winner != -1;
}

/* GameState.removeUnit, GameState.java:79-82 */
static void remove_unit(OGame *g, int u) {
    int i = list_index_of(g, u);
    if (i >= 0) { memmove(&g->list[i], &g->list[i + 1], sizeof(int) * (g->n - i - 1)); g->n--; }
    int a = find_assign(g, u);
    if (a >= 0) { memmove(&g->asg[a], &g->asg[a + 1], sizeof(OAssign) * (g->na - a - 1)); g->na--; }
}

static OAct mk_act(int type, int param, int x, int y, int utype) {
    OAct a; memset(&a, 0, sizeof a);
    a.type = type; a.param = param; a.x = x; a.y = y; a.utype = utype;
    return a;
}
static OAct act_none(int duration) { return mk_act(O_NONE, duration, 0, 0, -1); }
static OAct act_from_v(const OActionV *v) { return mk_act(v->type, v->param, v->x, v->y, v->utype); }
static OActionV act_to_v(const OAct *a) { OActionV v = {a->type, a->param, a->x, a->y, a->utype}; return v; }

/* UnitAction.resourceUsage, UnitAction.java:246-296 (cached on first call) */
static const ORu1 *act_ru(OAct *a, const OGame *g, int u) {
    if (a->ru_done) return &a->ru;
    a->ru_done = 1;
    a->ru.npos = 0; a->ru.pos = 0; a->ru.res[0] = a->ru.res[1] = 0;
    const OUnit *un = &g->pool[u];
    if (a->type == O_MOVE || a->type == O_PRODUCE) {
        if (a->type == O_PRODUCE) a->ru.res[un->player] += g->utt->f[a->utype][OF_COST];
        int pos = un->x + un->y * g->w;
        switch (a->param) {
            case O_UP: pos -= g->w; break;
            case O_RIGHT: pos++; break;
            case O_DOWN: pos += g->w; break;
            case O_LEFT: pos--; break;
        }
        a->ru.npos = 1; a->ru.pos = pos;
    }
    return &a->ru;
}

/* UnitAction.ETA, UnitAction.java:307-329 */
static int act_eta(const OAct *a, const OGame *g, int u) {
    const int16_t *f = g->utt->f[g->pool[u].type];
    switch (a->type) {
        case O_NONE: return a->param;
        case O_MOVE: return f[OF_MOVE_T];
        case O_ATTACK: return f[OF_ATTACK_T];
        case O_HARVEST: return f[OF_HARVEST_T];
        case O_RETURN: return f[OF_MOVE_T]; /* sic: moveTime */
        case O_PRODUCE: return g->utt->f[a->utype][OF_PRODUCE_T];
    }
    return 0;
}

/* UnitAction.equals, UnitAction.java:192-208 */
static int act_equals(const OAct *a, const OAct *b) {
    if (a->type != b->type) return 0;
    if (a->type == O_NONE || a->type == O_MOVE || a->type == O_HARVEST || a->type == O_RETURN) return a->param == b->param;
    if (a->type == O_ATTACK) return a->x == b->x && a->y == b->y;
    return a->param == b->param && a->utype == b->utype;
}

static int terrain_at(const OGame *g, int x, int y) { return g->terrain[x + y * g->w]; }

/* Unit.getUnitActions(s, noneDuration), Unit.java:382-522 */
static int unit_actions(const OGame *g, int u, int none_duration, OAct *l, int max) {
    const OUnit *me = &g->pool[u];
    const OUtt *t = g->utt;
    int fl = t->flags[me->type];
    int x = me->x, y = me->y, player = me->player;
    int n = 0;
#define PUSH(A) do { if (n < max) l[n] = (A); n++; } while (0)
    int uup = -1, uright = -1, udown = -1, uleft = -1;
    for (int i = 0; i < g->n; i++) {
        const OUnit *o = &g->pool[g->list[i]];
        if (o->x == x) {
            if (o->y == y - 1) uup = g->list[i];
            else if (o->y == y + 1) udown = g->list[i];
        } else if (o->y == y) {
            if (o->x == x - 1) uleft = g->list[i];
            else if (o->x == x + 1) uright = g->list[i];
        }
    }
    const OUnit *P = g->pool;
    if (fl & OFL_ATTACK) {
        if (t->f[me->type][OF_RANGE] == 1) {
            if (y > 0 && uup >= 0 && P[uup].player != player && P[uup].player >= 0) PUSH(mk_act(O_ATTACK, -1, P[uup].x, P[uup].y, -1));
            if (x < g->w - 1 && uright >= 0 && P[uright].player != player && P[uright].player >= 0) PUSH(mk_act(O_ATTACK, -1, P[uright].x, P[uright].y, -1));
            if (y < g->h - 1 && udown >= 0 && P[udown].player != player && P[udown].player >= 0) PUSH(mk_act(O_ATTACK, -1, P[udown].x, P[udown].y, -1));
            if (x > 0 && uleft >= 0 && P[uleft].player != player && P[uleft].player >= 0) PUSH(mk_act(O_ATTACK, -1, P[uleft].x, P[uleft].y, -1));
        } else {
            int sq = t->f[me->type][OF_RANGE] * t->f[me->type][OF_RANGE];
            for (int i = 0; i < g->n; i++) {
                const OUnit *o = &P[g->list[i]];
                if (o->player < 0 || o->player == player) continue;
                int sdx = (o->x - x) * (o->x - x), sdy = (o->y - y) * (o->y - y);
                if (sdx + sdy <= sq) PUSH(mk_act(O_ATTACK, -1, o->x, o->y, -1));
            }
        }
    }
    if (fl & OFL_HARVEST) {
        if (me->res == 0) {
            if (y > 0 && uup >= 0 && (t->flags[P[uup].type] & OFL_RESOURCE)) PUSH(mk_act(O_HARVEST, O_UP, 0, 0, -1));
            if (x < g->w - 1 && uright >= 0 && (t->flags[P[uright].type] & OFL_RESOURCE)) PUSH(mk_act(O_HARVEST, O_RIGHT, 0, 0, -1));
            if (y < g->h - 1 && udown >= 0 && (t->flags[P[udown].type] & OFL_RESOURCE)) PUSH(mk_act(O_HARVEST, O_DOWN, 0, 0, -1));
            if (x > 0 && uleft >= 0 && (t->flags[P[uleft].type] & OFL_RESOURCE)) PUSH(mk_act(O_HARVEST, O_LEFT, 0, 0, -1));
        }
        if (me->res > 0) {
            if (y > 0 && uup >= 0 && (t->flags[P[uup].type] & OFL_STOCKPILE) && P[uup].player == player) PUSH(mk_act(O_RETURN, O_UP, 0, 0, -1));
            if (x < g->w - 1 && uright >= 0 && (t->flags[P[uright].type] & OFL_STOCKPILE) && P[uright].player == player) PUSH(mk_act(O_RETURN, O_RIGHT, 0, 0, -1));
            if (y < g->h - 1 && udown >= 0 && (t->flags[P[udown].type] & OFL_STOCKPILE) && P[udown].player == player) PUSH(mk_act(O_RETURN, O_DOWN, 0, 0, -1));
            if (x > 0 && uleft >= 0 && (t->flags[P[uleft].type] & OFL_STOCKPILE) && P[uleft].player == player) PUSH(mk_act(O_RETURN, O_LEFT, 0, 0, -1));
        }
    }
    int tup = (y > 0 ? terrain_at(g, x, y - 1) : 1);
    int tright = (x < g->w - 1 ? terrain_at(g, x + 1, y) : 1);
    int tdown = (y < g->h - 1 ? terrain_at(g, x, y + 1) : 1);
    int tleft = (x > 0 ? terrain_at(g, x - 1, y) : 1);
    for (int k = 0; k < t->nprod[me->type]; k++) {
        int ut = t->prod[me->type][k];
        if (player >= 0 && g->res[player] >= t->f[ut][OF_COST]) {
            if (tup == 0 && unit_at(g, x, y - 1) < 0) PUSH(mk_act(O_PRODUCE, O_UP, 0, 0, ut));
            if (tright == 0 && unit_at(g, x + 1, y) < 0) PUSH(mk_act(O_PRODUCE, O_RIGHT, 0, 0, ut));
            if (tdown == 0 && unit_at(g, x, y + 1) < 0) PUSH(mk_act(O_PRODUCE, O_DOWN, 0, 0, ut));
            if (tleft == 0 && unit_at(g, x - 1, y) < 0) PUSH(mk_act(O_PRODUCE, O_LEFT, 0, 0, ut));
        }
    }
    if (fl & OFL_MOVE) {
        if (tup == 0 && uup < 0) PUSH(mk_act(O_MOVE, O_UP, 0, 0, -1));
        if (tright == 0 && uright < 0) PUSH(mk_act(O_MOVE, O_RIGHT, 0, 0, -1));
        if (tdown == 0 && udown < 0) PUSH(mk_act(O_MOVE, O_DOWN, 0, 0, -1));
        if (tleft == 0 && uleft < 0) PUSH(mk_act(O_MOVE, O_LEFT, 0, 0, -1));
    }
    PUSH(act_none(none_duration));
#undef PUSH
    return n;
}

#define MAX_UA 4200 /* ranged units can list one attack per enemy */

int o_unit_actions(const OGame *g, int unit_idx, int none_duration, OActionV *out, int max_out) {
    OAct *l = (OAct *)malloc(sizeof(OAct) * MAX_UA);
    int n = unit_actions(g, g->list[unit_idx], none_duration, l, MAX_UA);
    for (int i = 0; i < n && i < max_out; i++) out[i] = act_to_v(&l[i]);
    free(l);
    return n;
}

/* Unit.canExecuteAction, Unit.java:531-534 */
static int can_execute(const OGame *g, int u, const OAct *a) {
    if (a->type == O_PRODUCE && (a->utype < 0 || a->utype >= g->utt->n)) return 0; /* Java: NullPointerException */
    if (a->type < 0 || a->type > O_ATTACK) return 0;
    OAct *l = (OAct *)malloc(sizeof(OAct) * MAX_UA);
    int n = unit_actions(g, u, act_eta(a, g, u), l, MAX_UA);
    int ok = 0;
    for (int i = 0; i < n && !ok; i++) ok = act_equals(&l[i], a);
    free(l);
    return ok;
}

/* GameState.free, GameState.java:191-207 */
static int gs_free(const OGame *g, int x, int y) {
    if (terrain_at(g, x, y) != 0) return 0;
    for (int i = 0; i < g->n; i++) { const OUnit *u = &g->pool[g->list[i]]; if (u->x == x && u->y == y) return 0; }
    for (int i = 0; i < g->na; i++) {
        const OAssign *s = &g->asg[i];
        if (s->act.type == O_MOVE || s->act.type == O_PRODUCE) {
            const OUnit *u = &g->pool[s->unit];
            int d = s->act.param;
            if (d == O_UP && u->x == x && u->y == y + 1) return 0;
            if (d == O_RIGHT && u->x == x - 1 && u->y == y) return 0;
            if (d == O_DOWN && u->x == x && u->y == y - 1) return 0;
            if (d == O_LEFT && u->x == x + 1 && u->y == y) return 0;
        }
    }
    return 1;
}
int o_game_free_cell(const OGame *g, int x, int y) { return gs_free(g, x, y); }

/* GameState.getResourceUsage, GameState.java:652-664 (units in list order) */
static void gs_resource_usage(OGame *g, ORu *out) {
    for (int i = 0; i < g->n; i++) {
        int a = find_assign(g, g->list[i]);
        if (a >= 0) ru_merge1(out, act_ru(&g->asg[a].act, g, g->asg[a].unit));
    }
}

/* GameState.issue, GameState.java:249-328 */
typedef struct { int unit; OAct act; } OPair;

static int gs_issue(OGame *g, int n, OPair *pa) {
    int ret = 0;
    for (int k = 0; k < n; k++) {
        int u = pa[k].unit;
        OAct a = pa[k].act;
        ORu1 ru = *act_ru(&a, g, u); /* :262, computed once */
        for (int j = 0; j < g->na; j++) { /* :263 */
            OAssign *e = &g->asg[j];
            if (!ru1_consistent_with_ru1(act_ru(&e->act, g, e->unit), &ru, g)) {
                if (e->time == g->time) {
                    int cancel_old = 0, cancel_new = 0;
                    switch (g->utt->conflict) {
                        default:
                        case 1: cancel_old = cancel_new = 1; break;
                        case 2: if (o_jr_next_int_bound(&g->rng_conflict, 2) == 0) cancel_new = 1; else cancel_old = 1; break;
                        case 3: if ((g->cancel_ctr % 2) == 0) cancel_new = 1; else cancel_old = 1; g->cancel_ctr++; break;
                    }
                    int d1 = act_eta(&e->act, g, e->unit);
                    int d2 = act_eta(&a, g, u);
                    int d = d1 < d2 ? d1 : d2;
                    if (cancel_old) e->act = act_none(d);
                    if (cancel_new) a = act_none(d);
                } else {
                    g->errors |= OE_INCONSISTENT_OLDER;
                    a = mk_act(O_NONE, -1, 0, 0, -1); /* :316 new UnitAction(TYPE_NONE): parameter stays -1 */
                }
            }
        }
        pa[k].act = a;
        /* :321-322 unitActions.put: existing key keeps its slot */
        int s = find_assign(g, u);
        if (s < 0) {
            if (g->na == g->asg_cap) { g->asg_cap = g->asg_cap ? g->asg_cap * 2 : 64; g->asg = (OAssign *)realloc(g->asg, sizeof(OAssign) * g->asg_cap); }
            s = g->na++;
        }
        g->asg[s].unit = u; g->asg[s].act = a; g->asg[s].time = g->time;
        act_ru(&g->asg[s].act, g, u);
        if (a.type != O_NONE) ret = 1;
    }
    return ret;
}

/* GameState.issueSafe, GameState.java:338-408 */
static int gs_issue_safe(OGame *g, int n, OPair *pa) {
    /* PlayerAction.integrityCheck, PlayerAction.java:244-259 */
    int player = -1;
    for (int k = 0; k < n; k++) {
        int p = g->pool[pa[k].unit].player;
        if (player == -1) player = p; else if (player != p) { g->errors |= OE_MIXED_OWNERS; return -1; }
    }
    for (int k = 0; k < n; k++) {
        int u = pa[k].unit;
        if (!can_execute(g, u, &pa[k].act)) { /* :347-354 */
            int l = (pa[k].act.type == O_PRODUCE && (pa[k].act.utype < 0 || pa[k].act.utype >= g->utt->n)) ? 0 : act_eta(&pa[k].act, g, u);
            pa[k].act = act_none(l);
        }
        const ORu1 *r = act_ru(&pa[k].act, g, u); /* :386-399 */
        if (r->npos) {
            int y = r->pos / g->w, x = r->pos % g->w;
            if (terrain_at(g, x, y) != 0 || unit_at(g, x, y) >= 0) pa[k].act = act_none(act_eta(&pa[k].act, g, u));
        }
    }
    return gs_issue(g, n, pa);
}

int o_game_issue(OGame *g, int n, const int32_t *unit_idx, const OActionV *acts, int safe) {
    OPair *pa = (OPair *)malloc(sizeof(OPair) * (n > 0 ? n : 1));
    for (int k = 0; k < n; k++) {
        if (unit_idx[k] < 0 || unit_idx[k] >= g->n) { g->errors |= OE_BAD_UNIT; free(pa); return -1; }
        pa[k].unit = g->list[unit_idx[k]];
        pa[k].act = act_from_v(&acts[k]);
    }
    int r = safe ? gs_issue_safe(g, n, pa) : gs_issue(g, n, pa);
    free(pa);
    return r;
}

/* issue / issueSafe that hands the PlayerAction back as the call left it: issueSafe replaces illegal actions by NONE inside the
 * caller's PlayerAction (GameState.java:347-354,386-399), which is what the reward functions later read from the TraceEntry */
int o_game_issue_out(OGame *g, int n, const int32_t *unit_idx, OActionV *acts, int safe) {
    OPair *pa = (OPair *)malloc(sizeof(OPair) * (n > 0 ? n : 1));
    for (int k = 0; k < n; k++) {
        if (unit_idx[k] < 0 || unit_idx[k] >= g->n) { g->errors |= OE_BAD_UNIT; free(pa); return -1; }
        pa[k].unit = g->list[unit_idx[k]];
        pa[k].act = act_from_v(&acts[k]);
    }
    int r = safe ? gs_issue_safe(g, n, pa) : gs_issue(g, n, pa);
    for (int k = 0; k < n; k++) acts[k] = act_to_v(&pa[k].act);
    free(pa);
    return r;
}

/* UnitAction.execute, UnitAction.java:338-465 */
static int neighbour(const OGame *g, const OUnit *u, int dir) {
    switch (dir) {
        case O_UP: return unit_at(g, u->x, u->y - 1);
        case O_RIGHT: return unit_at(g, u->x + 1, u->y);
        case O_DOWN: return unit_at(g, u->x, u->y + 1);
        case O_LEFT: return unit_at(g, u->x - 1, u->y);
    }
    return -1;
}
static void act_execute(OGame *g, const OAct *a, int ui) {
    OUnit *u = &g->pool[ui];
    const OUtt *t = g->utt;
    switch (a->type) {
        case O_NONE: break;
        case O_MOVE:
            switch (a->param) {
                case O_UP: u->y--; break;
                case O_RIGHT: u->x++; break;
                case O_DOWN: u->y++; break;
                case O_LEFT: u->x--; break;
            }
            break;
        case O_ATTACK: {
            int o = unit_at(g, a->x, a->y);
            if (o >= 0) {
                int mn = t->f[u->type][OF_MINDMG], mx = t->f[u->type][OF_MAXDMG];
                int dmg = (mn == mx) ? mn : mn + o_jr_next_int_bound(&g->rng_damage, 1 + (mx - mn));
                g->pool[o].hp -= dmg;
                if (g->pool[o].hp <= 0) remove_unit(g, o);
            }
        } break;
        case O_HARVEST: {
            int r = neighbour(g, u, a->param);
            if (r >= 0 && (t->flags[g->pool[r].type] & OFL_RESOURCE) && (t->flags[u->type] & OFL_HARVEST) && u->res == 0) {
                int amt = t->f[u->type][OF_HARVEST_AMT];
                g->pool[r].res -= amt;
                if (g->pool[r].res <= 0) remove_unit(g, r);
                u = &g->pool[ui];
                u->res = amt;
            }
        } break;
        case O_RETURN: {
            int b = neighbour(g, u, a->param);
            if (b >= 0 && (t->flags[g->pool[b].type] & OFL_STOCKPILE) && u->res > 0) {
                g->res[u->player] += u->res;
                u->res = 0;
            }
        } break;
        case O_PRODUCE: {
            int tx = u->x, ty = u->y;
            switch (a->param) {
                case O_UP: ty--; break;
                case O_RIGHT: tx++; break;
                case O_DOWN: ty++; break;
                case O_LEFT: tx--; break;
            }
            int player = u->player, ut = a->utype;
            /* new Unit(player,type,x,y,0): ID = next_ID++ even if the unit is then not added (Unit.java:95-103) */
            int64_t id = g->next_id++;
            int cost = t->f[ut][OF_COST];
            if (g->res[player] - cost >= 0) {
                int nu = pool_new(g);
                OUnit *p = &g->pool[nu];
                p->type = ut; p->player = player; p->x = tx; p->y = ty; p->res = 0; p->hp = t->f[ut][OF_HP]; p->id = id;
                if (add_unit(g, nu)) g->res[player] -= cost;
            } else {
                g->errors |= OE_FAILED_PRODUCE;
            }
        } break;
    }
}

/* GameState.cycle, GameState.java:553-571 */
int o_game_cycle(OGame *g) {
    g->time++;
    int nr = 0;
    OAssign *ready = (OAssign *)malloc(sizeof(OAssign) * (g->na > 0 ? g->na : 1));
    for (int i = 0; i < g->na; i++)
        if (act_eta(&g->asg[i].act, g, g->asg[i].unit) + g->asg[i].time <= g->time) ready[nr++] = g->asg[i];
    for (int i = 0; i < nr; i++) {
        int a = find_assign(g, ready[i].unit); /* unitActions.remove(uaa.unit) */
        if (a >= 0) { memmove(&g->asg[a], &g->asg[a + 1], sizeof(OAssign) * (g->na - a - 1)); g->na--; }
        act_execute(g, &ready[i].act, ready[i].unit); /* runs even if the unit died earlier in this loop */
    }
    free(ready);
    return o_game_gameover(g);
}

/* GameState.isComplete, GameState.java:148-157 */
int o_game_is_complete(const OGame *g) {
    for (int i = 0; i < g->n; i++)
        if (g->pool[g->list[i]].player != -1 && find_assign(g, g->list[i]) < 0) return 0;
    return 1;
}

/* GameState.canExecuteAnyAction, GameState.java:416-423 */
static int can_execute_any(const OGame *g, int p) {
    for (int i = 0; i < g->n; i++)
        if (g->pool[g->list[i]].player == p && find_assign(g, g->list[i]) < 0) return 1;
    return 0;
}

/* GameState.getNextChangeTime, GameState.java:532-546 */
int o_game_next_change_time(const OGame *g) {
    int next = -1;
    for (int p = 0; p < 2; p++) if (can_execute_any(g, p)) return g->time;
    for (int i = 0; i < g->na; i++) {
        int t = g->asg[i].time + act_eta(&g->asg[i].act, g, g->asg[i].unit);
        if (next == -1 || t < next) next = t;
    }
    if (next == -1) return g->time;
    return next;
}

/* GameState.isUnitActionAllowed, GameState.java:434-457 */
static int is_unit_action_allowed(OGame *g, int u, OAct *ua) {
    if (ua->type == O_MOVE) {
        int x2 = g->pool[u].x + DX[ua->param], y2 = g->pool[u].y + DY[ua->param];
        if (x2 < 0 || y2 < 0 || x2 >= g->w || y2 >= g->h || terrain_at(g, x2, y2) == 1 || unit_at(g, x2, y2) >= 0) return 0;
    }
    ORu r; ru_init(&r);
    gs_resource_usage(g, &r);
    int ok = ru1_consistent_with_ru(act_ru(ua, g, u), &r, g);
    ru_free(&r);
    return ok;
}

/* ------------------------------------------------------------------------------------------------
 * RandomBiasedAI.getAction (ai/RandomBiasedAI.java:51-107) + Sampler.weighted (util/Sampler.java:116-137)
 * ---------------------------------------------------------------------------------------------- */
static int rb_get_action(OGame *g, int player, OPair *out) {
    int n = 0;
    if (!can_execute_any(g, player)) return 0;
    ORu par; ru_init(&par);
    gs_resource_usage(g, &par); /* :59-65, units in list order */
    OAct *l = (OAct *)malloc(sizeof(OAct) * MAX_UA);
    for (int i = 0; i < g->n; i++) {
        int u = g->list[i];
        if (g->pool[u].player != player || find_assign(g, u) >= 0) continue;
        int na = unit_actions(g, u, 10, l, MAX_UA);
        int none = -1;
        double total = 0, accum = 0, tmp;
        for (int k = 0; k < na; k++) {
            if (l[k].type == O_NONE) none = k;
            total += (l[k].type == O_ATTACK || l[k].type == O_HARVEST || l[k].type == O_RETURN) ? 5.0 : 1.0;
        }
        tmp = o_jr_next_double(&g->rng_policy) * total;
        int pick = -1;
        for (int k = 0; k < na; k++) {
            accum += (l[k].type == O_ATTACK || l[k].type == O_HARVEST || l[k].type == O_RETURN) ? 5.0 : 1.0;
            if (accum >= tmp) { pick = k; break; }
        }
        if (pick < 0) pick = none; /* Sampler throws -> catch -> none */
        OAct ua = l[pick];
        if (ru1_consistent_with_ru(act_ru(&ua, g, u), &par, g)) {
            ru_merge1(&par, act_ru(&ua, g, u));
            out[n].unit = u; out[n].act = ua; n++;
        } else {
            out[n].unit = u; out[n].act = l[none]; n++;
        }
    }
    free(l);
    ru_free(&par);
    return n;
}

static int pairs_out(const OGame *g, int n, const OPair *pa, int32_t *unit_idx, OActionV *acts) {
    for (int k = 0; k < n; k++) { unit_idx[k] = list_index_of(g, pa[k].unit); acts[k] = act_to_v(&pa[k].act); }
    return n;
}

int o_ai_random_biased(OGame *g, int player, int32_t *unit_idx, OActionV *acts) {
    OPair *pa = (OPair *)malloc(sizeof(OPair) * (g->n + 1));
    int n = rb_get_action(g, player, pa);
    pairs_out(g, n, pa, unit_idx, acts);
    free(pa);
    return n;
}

/* ------------------------------------------------------------------------------------------------
 * Pathfinding: AStarPathFinding.java:52-79,104-138,175-295 ; BFSPathFinding.java:41-147
 * ---------------------------------------------------------------------------------------------- */
typedef struct { int n; int8_t *free_; int *closed, *open, *heur, *parents, *cost, *inoc; } OPf;

static void pf_alloc(OPf *p, int n) {
    p->n = n;
    p->free_ = (int8_t *)malloc(n);
    p->closed = (int *)malloc(sizeof(int) * n * 6);
    p->open = p->closed + n; p->heur = p->open + n; p->parents = p->heur + n; p->cost = p->parents + n; p->inoc = p->cost + n;
}
static void pf_release(OPf *p) { free(p->free_); free(p->closed); }
static int manh(int x, int y, int x2, int y2) { return abs(x - x2) + abs(y - y2); }

static int pf_free(OPf *p, const OGame *g, int x, int y) {
    int8_t *c = &p->free_[x + y * g->w];
    if (*c < 0) *c = (int8_t)gs_free(g, x, y);
    return *c;
}

/* AStarPathFinding.addToOpen :104-138; returns updated openinsert */
static int astar_add(OPf *p, int openinsert, int newPos, int oldPos, int h) {
    p->cost[newPos] = p->cost[oldPos] + 1;
    int at = 0;
    for (int i = openinsert - 1; i >= 0; i--) {
        if (p->heur[i] + p->cost[p->open[i]] >= h + p->cost[newPos]) { at = i + 1; break; }
    }
    memmove(&p->open[at + 1], &p->open[at], sizeof(int) * (openinsert - at));
    memmove(&p->heur[at + 1], &p->heur[at], sizeof(int) * (openinsert - at));
    memmove(&p->parents[at + 1], &p->parents[at], sizeof(int) * (openinsert - at));
    p->open[at] = newPos; p->heur[at] = h; p->parents[at] = oldPos;
    p->inoc[newPos] = 1;
    return openinsert + 1;
}

static int decode_first_step(const OPf *p, int pos, int parent, int w) {
    int last = pos;
    while (parent != pos) { last = pos; pos = parent; parent = p->closed[pos]; }
    if (last == pos + w) return O_DOWN;
    if (last == pos - 1) return O_LEFT;
    if (last == pos - w) return O_UP;
    if (last == pos + 1) return O_RIGHT;
    return -1;
}

/* GreedyPathFinding.findPathToPositionInRange, GreedyPathFinding.java:53-84: the free neighbour closest (Euclidean^2) to the
 * target; the first free direction is always taken, a later one only if strictly closer; "already in range" compares the
 * SQUARED distance with the unsquared range (:66), as the reference does */
static int pf_greedy(const OGame *g, int start, int targetpos, int range, const ORu *ru) {
    int w = g->w, h = g->h;
    const OUnit *s = &g->pool[start];
    int x1 = s->x, y1 = s->y, x2 = targetpos % w, y2 = targetpos / w;
    int min_d = (x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1), direction = -1;
    if (range >= 0 && min_d <= range) return -1; /* range < 0: findPath (:16-47), which has no such test */
    for (int i = 0; i < 4; i++) {
        int x = x1 + DX[i], y = y1 + DY[i];
        if (x >= 0 && x < w && y >= 0 && y < h && gs_free(g, x, y)) {
            if (ru && ru_contains(ru, x + y * w)) continue;
            int d = (x2 - x) * (x2 - x) + (y2 - y) * (y2 - y);
            if (direction == -1 || d < min_d) { min_d = d; direction = i; }
        }
    }
    return direction;
}

static int pf_find(const OGame *g, int kind, int start, int targetpos, int range, const ORu *ru) {
    int w = g->w, h = g->h, n = w * h;
    if (kind == O_PF_GREEDY) return pf_greedy(g, start, targetpos, range, ru);
    if (range < 0) range = 0; /* AStarPathFinding.findPath :43-45, BFSPathFinding.findPath :33-35 */
    OPf p; pf_alloc(&p, n);
    memset(p.free_, -1, n);
    for (int i = 0; i < n; i++) { p.closed[i] = -1; p.inoc[i] = 0; }
    if (ru) for (int i = 0; i < ru->npos; i++) { int q = ru->pos[i]; if (q >= 0 && q < n) p.free_[q] = 0; }
    int tx = targetpos % w, ty = targetpos / w;
    int sq = range * range;
    const OUnit *s = &g->pool[start];
    int startPos = s->y * w + s->x;
    int result = -1;
    if (kind == O_PF_ASTAR) {
        int oi = 0;
        p.open[0] = startPos; p.heur[0] = manh(s->x, s->y, tx, ty); p.parents[0] = startPos; p.inoc[startPos] = 1; p.cost[startPos] = 0; oi = 1;
        while (oi > 0) {
            oi--;
            int pos = p.open[oi], parent = p.parents[oi];
            if (p.closed[pos] != -1) continue;
            p.closed[pos] = parent;
            int x = pos % w, y = pos / w;
            if ((x - tx) * (x - tx) + (y - ty) * (y - ty) <= sq) { result = decode_first_step(&p, pos, parent, w); break; }
            if (y > 0 && p.inoc[pos - w] == 0 && pf_free(&p, g, x, y - 1)) oi = astar_add(&p, oi, pos - w, pos, manh(x, y - 1, tx, ty));
            if (x < w - 1 && p.inoc[pos + 1] == 0 && pf_free(&p, g, x + 1, y)) oi = astar_add(&p, oi, pos + 1, pos, manh(x + 1, y, tx, ty));
            if (y < h - 1 && p.inoc[pos + w] == 0 && pf_free(&p, g, x, y + 1)) oi = astar_add(&p, oi, pos + w, pos, manh(x, y + 1, tx, ty));
            if (x > 0 && p.inoc[pos - 1] == 0 && pf_free(&p, g, x - 1, y)) oi = astar_add(&p, oi, pos - 1, pos, manh(x - 1, y, tx, ty));
        }
    } else {
        int oi = 0, orm = 0;
        p.open[0] = startPos; p.parents[0] = startPos; p.inoc[startPos] = 1; oi = 1;
        while (oi != orm) {
            int pos = p.open[orm], parent = p.parents[orm];
            orm++; if (orm >= n) orm = 0;
            if (p.closed[pos] != -1) continue;
            p.closed[pos] = parent;
            int x = pos % w, y = pos / w;
            if ((x - tx) * (x - tx) + (y - ty) * (y - ty) <= sq) { result = decode_first_step(&p, pos, parent, w); break; }
#define BFS_PUSH(NP) do { p.open[oi] = (NP); p.parents[oi] = pos; oi++; if (oi >= n) oi = 0; p.inoc[(NP)] = 1; } while (0)
            if (y > 0 && p.inoc[pos - w] == 0 && pf_free(&p, g, x, y - 1)) BFS_PUSH(pos - w);
            if (x < w - 1 && p.inoc[pos + 1] == 0 && pf_free(&p, g, x + 1, y)) BFS_PUSH(pos + 1);
            if (y < h - 1 && p.inoc[pos + w] == 0 && pf_free(&p, g, x, y + 1)) BFS_PUSH(pos + w);
            if (x > 0 && p.inoc[pos - 1] == 0 && pf_free(&p, g, x - 1, y)) BFS_PUSH(pos - 1);
#undef BFS_PUSH
        }
    }
    pf_release(&p);
    return result;
}

/* FloodFillPathFinding, ai/abstraction/pathfinding/FloodFillPathFinding.java (whole file): an instance keeps `cache`, a map from target
 * position to a distance array, and `lastFrame`.  PARITY UNPINNED (no golden data in the reference). */
typedef struct { int w, h; int **cache; /* [w*h] -> int[w*h] (distances[x][y] at x*h + y) or NULL */ int last_frame; } OFf;
static OFf *ff_new(void) { OFf *f = (OFf *)calloc(1, sizeof(OFf)); f->last_frame = -1; return f; }
static void ff_clear(OFf *f) { if (f->cache) for (int i = 0; i < f->w * f->h; i++) { free(f->cache[i]); f->cache[i] = NULL; } }
static void ff_free(OFf *f) { if (!f) return; ff_clear(f); free(f->cache); free(f); }
/* getAction :139-166 */
static int ff_get_action(const OFf *f, const int *distances, int x, int y) {
    int w = f->w, h = f->h;
#define FF_B(X, Y) ((X) >= 0 && (Y) >= 0 && (X) < w && (Y) < h)
    int dists[4] = {FF_B(x - 1, y) ? distances[(x - 1) * h + y] : INT32_MAX, FF_B(x, y - 1) ? distances[x * h + y - 1] : INT32_MAX,
                    FF_B(x + 1, y) ? distances[(x + 1) * h + y] : INT32_MAX, FF_B(x, y + 1) ? distances[x * h + y + 1] : INT32_MAX};
    int index = 0, min = dists[0];
    for (int i = 1; i < 4; i++) if (dists[i] < min) { index = i; min = dists[i]; }
    if (min == INT32_MAX) return -1;
    switch (index) { case 0: return O_LEFT; case 1: return O_UP; case 2: return O_RIGHT; default: return O_DOWN; }
}
/* GameState.getAllFree :215-229 over PhysicalGameState.getAllFree :512-525 */
static void gs_all_free(const OGame *g, uint8_t *fr /* [x*h + y] */) {
    int w = g->w, h = g->h;
    for (int x = 0; x < w; x++) for (int y = 0; y < h; y++) fr[x * h + y] = terrain_at(g, x, y) == 0;
    for (int i = 0; i < g->n; i++) { const OUnit *u = &g->pool[g->list[i]]; fr[u->x * h + u->y] = 0; }
    for (int i = 0; i < g->na; i++) {
        const OAssign *a = &g->asg[i];
        if (a->act.type == O_MOVE || a->act.type == O_PRODUCE) {
            const OUnit *u = &g->pool[a->unit];
            int d = a->act.param, x = u->x + (d == O_RIGHT) - (d == O_LEFT), y = u->y + (d == O_DOWN) - (d == O_UP);
            if (d >= 0 && d < 4 && x >= 0 && y >= 0 && x < w && y < h) fr[x * h + y] = 0; /* (the reference would throw outside the map) */
        }
    }
}
static int pf_floodfill(OFf *f, const OGame *g, int start, int targetpos, int range, const ORu *ru) {
    const OUnit *s = &g->pool[start];
    int w = g->w, h = g->h;
    if (targetpos < 0 || targetpos >= w * h) return -1; /* the reference indexes distances[x][y] with the target (an exception off the map): null */
    if (range < 0) range = 0; /* findPath :38-40 */
    if (f->w != w || f->h != h) { ff_clear(f); free(f->cache); f->w = w; f->h = h; f->cache = (int **)calloc((size_t)w * h, sizeof(int *)); }
    int x = targetpos % w, y = targetpos / w;
    if ((s->x - x) * (s->x - x) + (s->y - y) * (s->y - y) <= range * range) return -1; /* already there */
    if (g->time < f->last_frame) ff_clear(f); /* new game */
    f->last_frame = g->time;
    uint8_t *fre = (uint8_t *)malloc((size_t)w * h); memset(fre, 1, (size_t)w * h); /* initFree :127-138 */
    if (ru) for (int i = 0; i < ru->npos; i++) { int q = ru->pos[i]; if (q >= 0 && q < w * h) fre[(q % w) * h + q / w] = 0; }
    int result;
    if (f->cache[targetpos]) {
        int action = ff_get_action(f, f->cache[targetpos], s->x, s->y);
        int ok = 0;
        if (action >= 0) {
            int px = s->x + DX[action], py = s->y + DY[action];
            ok = fre[px * h + py] && gs_free(g, px, py);
        }
        if (ok) { free(fre); return action; }
        free(f->cache[targetpos]); f->cache[targetpos] = NULL; /* cache.remove(targetpos) */
    }
    /* calculateDistances :108-126 (ALT_THRESHOLD = 0: the alternative path finder is unreachable past the range test above) */
    int *distances = (int *)malloc(sizeof(int) * (size_t)w * h);
    for (int i = 0; i < w * h; i++) distances[i] = INT32_MAX;
    distances[x * h + y] = 0;
    { /* doFloodFill :47-107 */
        uint8_t *gsFree = (uint8_t *)malloc((size_t)w * h); gs_all_free(g, gsFree);
        int *fx = (int *)malloc(sizeof(int) * ((size_t)w * h + 8)), *fy = (int *)malloc(sizeof(int) * ((size_t)w * h + 8));
        int nf = 1, index = 0, reached = 0, finalX = s->x, finalY = s->y;
        fx[0] = x; fy[0] = y;
        static const int NX[4] = {-1, 0, 1, 0}, NY[4] = {0, -1, 0, 1}; /* left, up, right, down */
        while (index < nf) {
            int cx = fx[index], cy = fy[index];
            for (int k = 0; k < 4; k++) {
                int nx = cx + NX[k], ny = cy + NY[k];
                if (nx == finalX && ny == finalY) reached = 1;
                if (FF_B(nx, ny) && distances[nx * h + ny] == INT32_MAX && fre[nx * h + ny] && gsFree[nx * h + ny]) {
                    distances[nx * h + ny] = distances[cx * h + cy] + 1;
                    fx[nf] = nx; fy[nf] = ny; nf++;
                }
            }
            if (reached) break;
            index++;
        }
        free(gsFree); free(fx); free(fy);
    }
#undef FF_B
    f->cache[targetpos] = distances;
    result = ff_get_action(f, distances, s->x, s->y);
    free(fre);
    return result;
}

/* a FloodFillPathFinding instance of its own (known-answer tests): the cache lives as long as the handle */
void *o_ff_create(void) { return ff_new(); }
void o_ff_free(void *f) { ff_free((OFf *)f); }
int o_ff_find(void *f, const OGame *g, int unit_idx, int targetpos, int range, int n_ru, const int32_t *ru_pos) {
    ORu r; ru_init(&r);
    for (int i = 0; i < n_ru; i++) ru_add_pos(&r, ru_pos[i]);
    int d = pf_floodfill((OFf *)f, g, g->list[unit_idx], targetpos, range, &r);
    ru_free(&r);
    return d;
}

int o_pathfind(const OGame *g, int kind, int unit_idx, int targetpos, int range, int n_ru, const int32_t *ru_pos) {
    ORu r; ru_init(&r);
    for (int i = 0; i < n_ru; i++) ru_add_pos(&r, ru_pos[i]);
    int d = pf_find(g, kind, g->list[unit_idx], targetpos, range, &r);
    ru_free(&r);
    return d;
}

/* ------------------------------------------------------------------------------------------------
 * AbstractionLayerAI + WorkerRush + LightRush
 *   ai/abstraction/AbstractionLayerAI.java:58-113,143-245 ; WorkerRush.java:63-204 ; LightRush.java:77-258 ;
 *   Attack.java:51 ; Harvest.java:72 ; Build.java:54 ; Train.java:48-128
 * ---------------------------------------------------------------------------------------------- */
enum { AA_TRAIN = 1, AA_BUILD, AA_HARVEST, AA_ATTACK, AA_MOVE /* Move.java */, AA_RANGED_ATTACK /* cRush/RangedAttack.java: target + base = racks */,
       AA_TACTIC /* cRush/CRanged_Tactic.java: target, base = home, x = enemyBase (-1: null) */ };
typedef struct {
    int unit; int kind;
    int type;          /* train / build */
    int x, y;          /* build */
    int target, base;  /* harvest (target may be -1 = null) / attack (target) */
    int completed;     /* train */
} OAbs;

struct OAi { int kind; int pf; OAbs *a; int n, cap; int po_rush; OFf *ff; /* the instance's FloodFillPathFinding, pf == 3 */
             int building_racks, resources_used; /* CRush_V1.java:64-65 */ };
static int ai_pf(OAi *ai, const OGame *g, int start, int targetpos, int range, const ORu *ru) {
    if (ai->pf == 3) { if (!ai->ff) ai->ff = ff_new(); return pf_floodfill(ai->ff, g, start, targetpos, range, ru); }
    return pf_find(g, ai->pf, start, targetpos, range, ru);
}

OAi *o_ai_create(int kind, int pathfinder) {
    OAi *ai = (OAi *)calloc(1, sizeof(OAi));
    ai->kind = kind; ai->pf = pathfinder;
    return ai;
}
OAi *o_ai_clone(const OAi *s) { return o_ai_create(s->kind, s->pf); } /* AI.clone(): fresh instance, empty actions map */
void o_ai_free(OAi *ai) { if (ai) { ff_free(ai->ff); free(ai->a); free(ai); } }

static OAbs *ai_get(OAi *ai, int u) { for (int i = 0; i < ai->n; i++) if (ai->a[i].unit == u) return &ai->a[i]; return NULL; }
static void ai_put(OAi *ai, OAbs v) { /* LinkedHashMap.put: existing key keeps its slot */
    OAbs *e = ai_get(ai, v.unit);
    if (e) { *e = v; return; }
    if (ai->n == ai->cap) { ai->cap = ai->cap ? ai->cap * 2 : 32; ai->a = (OAbs *)realloc(ai->a, sizeof(OAbs) * ai->cap); }
    ai->a[ai->n++] = v;
}
static void ai_train(OAi *ai, int u, int type) { OAbs v = {u, AA_TRAIN, type, 0, 0, -1, -1, 0}; ai_put(ai, v); }
static void ai_build(OAi *ai, int u, int type, int x, int y) { OAbs v = {u, AA_BUILD, type, x, y, -1, -1, 0}; ai_put(ai, v); }
static void ai_harvest(OAi *ai, int u, int target, int base) { OAbs v = {u, AA_HARVEST, -1, 0, 0, target, base, 0}; ai_put(ai, v); }
static void ai_attack(OAi *ai, int u, int target) { OAbs v = {u, AA_ATTACK, -1, 0, 0, target, -1, 0}; ai_put(ai, v); }
static void ai_move(OAi *ai, int u, int x, int y) { OAbs v = {u, AA_MOVE, -1, x, y, -1, -1, 0}; ai_put(ai, v); }
static void ai_tactic(OAi *ai, int u, int target, int home, int eb) { OAbs v = {u, AA_TACTIC, -1, eb, 0, target, home, 0}; ai_put(ai, v); } /* CRush_V2.java:477-479 */
static void ai_ranged_attack(OAi *ai, int u, int target, int racks) { OAbs v = {u, AA_RANGED_ATTACK, -1, 0, 0, target, racks, 0}; ai_put(ai, v); } /* CRush_V1.java:419-421 */

/* unit types by name, as the reference looks them up (utt.getUnitType("Worker") ...): the ids of the standard tables */
static int type_by_role_base(void) { return 1; }
static int type_by_role_barracks(void) { return 2; }
static int type_by_role_worker(void) { return 3; }
static int type_by_role_light(void) { return 4; }

static int in_list(const OGame *g, int u) { return u >= 0 && list_index_of(g, u) >= 0; }

static int aa_completed(const OAbs *aa, const OGame *g) {
    switch (aa->kind) {
        case AA_TRAIN: return aa->completed;                                  /* Train.java:30 */
        case AA_BUILD: return unit_at(g, aa->x, aa->y) >= 0;                  /* Build.java:33-37 */
        case AA_HARVEST:                                                      /* Harvest.java:46-52 */
            if (g->pool[aa->unit].res > 0) return !in_list(g, aa->base);
            return !in_list(g, aa->target);
        case AA_ATTACK: return !in_list(g, aa->target);                       /* Attack.java:30-33 */
        case AA_RANGED_ATTACK: return !in_list(g, aa->target);                /* RangedAttack.java:36-39 */
        case AA_TACTIC: return !in_list(g, aa->target);                       /* CRanged_Tactic.java:59-62 */
        case AA_MOVE: return g->pool[aa->unit].x == aa->x && g->pool[aa->unit].y == aa->y; /* Move.java:29-31 */
    }
    return 1;
}

static int mk_move(OAct *out, int dir) { if (dir < 0) return 0; *out = mk_act(O_MOVE, dir, 0, 0, -1); return 1; }

/* Train.score, Train.java:98-126 */
static int train_score(const OGame *g, int x, int y, int type, int player) {
    int distance = 0, first = 1;
    for (int i = 0; i < g->n; i++) {
        const OUnit *u = &g->pool[g->list[i]];
        int ok = (g->utt->flags[type] & OFL_HARVEST) ? (g->utt->flags[u->type] & OFL_RESOURCE) != 0 : (u->player >= 0 && u->player != player);
        if (ok) { int d = abs(u->x - x) + abs(u->y - y); if (first || d < distance) { distance = d; first = 0; } }
    }
    return -distance;
}


/* cRush/CRanged_Tactic.java:77-307.  Distances are square roots of integers that are only compared with each other or with
 * integers, so their squares are compared instead. */
static int d2_units(const OUnit *a, const OUnit *b) { int dx = b->x - a->x, dy = b->y - a->y; return dx * dx + dy * dy; }
static int tactic_unit_at(const OGame *g, int x, int y) { /* PhysicalGameState.getUnitAt: null off the map */
    if (x < 0 || y < 0 || x >= g->w || y >= g->h) return -1;
    return unit_at(g, x, y);
}
static int tactic_allowed(OGame *g, int unit, int dir, OAct *out) {
    OAct mv;
    if (mk_move(&mv, dir) && is_unit_action_allowed(g, unit, &mv)) { *out = mv; return 1; }
    return 0;
}
static int tactic_execute(OAi *ai, OAbs *aa, OGame *g, const ORu *ru, OAct *out) {
    const OUtt *t = g->utt;
    const int WORKER = type_by_role_worker(), BASE = type_by_role_base(), LIGHT = type_by_role_light(), HEAVY = 5, RANGED = 6;
    const int w = g->w, u = aa->unit;
    const OUnit *unit = &g->pool[u], *target = &g->pool[aa->target];
    const int player = unit->player;
    const OUnit *home = aa->base >= 0 ? &g->pool[aa->base] : unit;            /* :84-86 */
    const OUnit *eb = aa->x >= 0 ? &g->pool[aa->x] : target;                  /* :88-90 */
    const int range = t->f[unit->type][OF_RANGE];
    const int rd2 = d2_units(unit, home), d2 = d2_units(unit, target);
    int n_enemy_bases = 0, enemy_attack_units = 0, enemy_workers = 0;
    const int cutoff = (g->w * g->h > 3000) ? 15000 : 5000;
    for (int i = 0; i < g->n; i++) { /* `player != p.getID()`: neutral units count as well, only their types never match */
        const OUnit *o = &g->pool[g->list[i]];
        if (o->player == player) continue;
        if (o->type == BASE) n_enemy_bases++;
        if (o->type == RANGED || o->type == HEAVY || o->type == LIGHT) enemy_attack_units++;
        if (o->type == WORKER) enemy_workers++;
    }
    int time_to_attack = ((enemy_workers < 2 * n_enemy_bases || n_enemy_bases == 0) && enemy_attack_units == 0) || g->time > cutoff;
    /* nearestRangedAlly(enemyBase) :367-389: first own Ranged unit with the smallest distance to the enemy base */
    int ally = -1, best = -1;
    for (int i = 0; i < g->n; i++) {
        const OUnit *o = &g->pool[g->list[i]];
        if (o->player == player && o->type == RANGED) { int dd = d2_units(o, eb); if (best == -1 || dd < best) { best = dd; ally = g->list[i]; } }
    }
    const int ad2 = ally >= 0 ? d2_units(&g->pool[ally], target) : 0;
    const int tpos = target->x + target->y * w;
    if (unit->type == WORKER) { /* :149-176 */
        if (d2 <= range * range) { *out = mk_act(O_ATTACK, -1, target->x, target->y, -1); return 1; }
        int dir;
        if (time_to_attack) dir = ai_pf(ai, g, u, tpos, range, ru);
        else if (ally >= 0) {
            const OUnit *al = &g->pool[ally];
            if (d2 > ad2) dir = ai_pf(ai, g, u, tpos, range, ru);
            else dir = ai_pf(ai, g, u, al->x + al->y * w, range, ru);
            if (dir < 0) dir = ai_pf(ai, g, u, (al->x - 1) + al->y * w, range + 1, ru);
            if (dir < 0) dir = ai_pf(ai, g, u, tpos, range, ru);
        } else dir = ai_pf(ai, g, u, tpos, range, ru);
        return tactic_allowed(g, u, dir, out);
    }
    if (d2 <= range * range) { *out = mk_act(O_ATTACK, -1, target->x, target->y, -1); return 1; } /* :179-181 */
    if (ally < 0 || ally == u) { /* the unit leads: :182-194 */
        int dir = -1;
        if (time_to_attack && target->type == BASE) dir = ai_pf(ai, g, u, tpos, range, ru);
        else if (rd2 < 25 || d2_units(unit, eb) > d2_units(home, eb)) dir = ai_pf(ai, g, u, eb->x + eb->y * w, range, ru);
        return tactic_allowed(g, u, dir, out);
    }
    if (time_to_attack) { /* :195-216 */
        if (range >= 1 && d2 <= (range - 1) * (range - 1) && rd2 > 4 && t->f[unit->type][OF_MOVE_T] < t->f[target->type][OF_MOVE_T])
            return tactic_allowed(g, u, ai_pf(ai, g, u, home->x + home->y * w, range, ru), out);
        if (d2 <= range * range) { *out = mk_act(O_ATTACK, -1, target->x, target->y, -1); return 1; }
        return tactic_allowed(g, u, ai_pf(ai, g, u, tpos, range, ru), out);
    }
    /* line up next to the leading ranged unit :218-268, squareMove :290-343 */
    int ax = g->pool[ally].x, ay = g->pool[ally].y;
#define D2XY(X, Y) (((eb->x) - (X)) * ((eb->x) - (X)) + ((eb->y) - (Y)) * ((eb->y) - (Y)))
    int sgn = D2XY(ax, ay + 1) > D2XY(ax, ay) ? 1 : -1; /* down / right of the leader, or up / left */
    for (;;) {
        int a = tactic_unit_at(g, ax, ay + sgn), b = tactic_unit_at(g, ax + sgn, ay + sgn), c = tactic_unit_at(g, ax + sgn, ay);
        int found = 0;
        if (a >= 0 && a != u && b >= 0 && b != u && c >= 0 && c != u) { ax = g->pool[c].x; ay = g->pool[c].y; } else found = 1;
        a = tactic_unit_at(g, ax, ay + sgn); b = tactic_unit_at(g, ax + sgn, ay + sgn); c = tactic_unit_at(g, ax + sgn, ay);
        if (a < 0 || b < 0 || c < 0) found = 1;
        if (found) break;
    }
    {
        int sg = D2XY(ax, ay + 1) > D2XY(ax, ay) ? 1 : -1; /* squareMove decides again, around the unit it was handed */
        int a = tactic_unit_at(g, ax, ay + sg), b = tactic_unit_at(g, ax + sg, ay + sg), c = tactic_unit_at(g, ax + sg, ay);
        if (u == a || u == b || u == c) return 0;
        int dir = -1; /* pf.findPath = range 0; linear positions, as the reference computes them (a column off the map aliases) */
        if (a < 0) dir = ai_pf(ai, g, u, ax + (ay + sg) * w, -1, ru);
        else if (c < 0) dir = ai_pf(ai, g, u, (ax + sg) + ay * w, -1, ru);
        else if (b < 0) dir = ai_pf(ai, g, u, (ax + sg) + (ay + sg) * w, -1, ru);
        return tactic_allowed(g, u, dir, out);
    }
#undef D2XY
}

/* returns 1 and fills *out if the abstract action yields a unit action, 0 for null */
static int aa_execute(OAi *ai, OAbs *aa, OGame *g, const ORu *ru, OAct *out) {
    const OUnit *unit = &g->pool[aa->unit];
    int w = g->w;
    switch (aa->kind) {
        case AA_ATTACK: { /* Attack.java:51-64 */
            const OUnit *t = &g->pool[aa->target];
            int dx = t->x - unit->x, dy = t->y - unit->y;
            double d = sqrt((double)(dx * dx + dy * dy));
            int range = g->utt->f[unit->type][OF_RANGE];
            if (d <= range) { *out = mk_act(O_ATTACK, -1, t->x, t->y, -1); return 1; }
            OAct mv;
            if (mk_move(&mv, ai_pf(ai, g, aa->unit, t->x + t->y * w, range, ru)) && is_unit_action_allowed(g, aa->unit, &mv)) { *out = mv; return 1; }
            return 0;
        }
        case AA_RANGED_ATTACK: { /* cRush/RangedAttack.java:58-87: step back towards the barracks while a slower enemy is well
                                  * inside the range, shoot when in range, approach otherwise */
            const OUnit *t = &g->pool[aa->target];
            int range = g->utt->f[unit->type][OF_RANGE];
            double rd = 0.0;
            if (aa->base >= 0) { const OUnit *r = &g->pool[aa->base]; int rdx = r->x - unit->x, rdy = r->y - unit->y; rd = sqrt((double)(rdx * rdx + rdy * rdy)); }
            int dx = t->x - unit->x, dy = t->y - unit->y;
            double d = sqrt((double)(dx * dx + dy * dy));
            OAct mv;
            if (d <= range - 1 && rd > 2 && g->utt->f[unit->type][OF_MOVE_T] < g->utt->f[t->type][OF_MOVE_T]) {
                const OUnit *r = &g->pool[aa->base];
                if (mk_move(&mv, ai_pf(ai, g, aa->unit, r->x + r->y * w, range, ru)) && is_unit_action_allowed(g, aa->unit, &mv)) { *out = mv; return 1; }
                return 0;
            }
            if (d <= range) { *out = mk_act(O_ATTACK, -1, t->x, t->y, -1); return 1; }
            if (mk_move(&mv, ai_pf(ai, g, aa->unit, t->x + t->y * w, range, ru)) && is_unit_action_allowed(g, aa->unit, &mv)) { *out = mv; return 1; }
            return 0;
        }
        case AA_TACTIC: return tactic_execute(ai, aa, g, ru, out);
        case AA_MOVE: { /* Move.java:49-55: pf.findPath */
            OAct mv;
            if (mk_move(&mv, ai_pf(ai, g, aa->unit, aa->x + aa->y * w, -1, ru)) && is_unit_action_allowed(g, aa->unit, &mv)) { *out = mv; return 1; }
            return 0;
        }
        case AA_HARVEST: { /* Harvest.java:72-113 */
            int other = (unit->res == 0) ? aa->target : aa->base;
            int atype = (unit->res == 0) ? O_HARVEST : O_RETURN;
            if (other < 0) return 0;
            const OUnit *t = &g->pool[other];
            OAct mv;
            if (mk_move(&mv, ai_pf(ai, g, aa->unit, t->x + t->y * w, 1, ru))) {
                if (is_unit_action_allowed(g, aa->unit, &mv)) { *out = mv; return 1; }
                return 0;
            }
            if (t->x == unit->x && t->y == unit->y - 1) { *out = mk_act(atype, O_UP, 0, 0, -1); return 1; }
            if (t->x == unit->x + 1 && t->y == unit->y) { *out = mk_act(atype, O_RIGHT, 0, 0, -1); return 1; }
            if (t->x == unit->x && t->y == unit->y + 1) { *out = mk_act(atype, O_DOWN, 0, 0, -1); return 1; }
            if (t->x == unit->x - 1 && t->y == unit->y) { *out = mk_act(atype, O_LEFT, 0, 0, -1); return 1; }
            return 0;
        }
        case AA_BUILD: { /* Build.java:54-77 */
            OAct mv;
            if (mk_move(&mv, ai_pf(ai, g, aa->unit, aa->x + aa->y * w, 1, ru))) {
                if (is_unit_action_allowed(g, aa->unit, &mv)) { *out = mv; return 1; }
                return 0;
            }
            int dir = -1;
            if (aa->x == unit->x && aa->y == unit->y - 1) dir = O_UP;
            if (aa->x == unit->x + 1 && aa->y == unit->y) dir = O_RIGHT;
            if (aa->x == unit->x && aa->y == unit->y + 1) dir = O_DOWN;
            if (aa->x == unit->x - 1 && aa->y == unit->y) dir = O_LEFT;
            if (dir >= 0) { OAct ua = mk_act(O_PRODUCE, dir, 0, 0, aa->type); if (is_unit_action_allowed(g, aa->unit, &ua)) { *out = ua; return 1; } }
            return 0;
        }
        case AA_TRAIN: { /* Train.java:48-95 */
            int x = unit->x, y = unit->y, best_dir = -1, best = -1;
            if (y > 0 && gs_free(g, x, y - 1)) { int s = train_score(g, x, y - 1, aa->type, unit->player); if (s > best || best_dir == -1) { best = s; best_dir = O_UP; } }
            if (x < g->w - 1 && gs_free(g, x + 1, y)) { int s = train_score(g, x + 1, y, aa->type, unit->player); if (s > best || best_dir == -1) { best = s; best_dir = O_RIGHT; } }
            if (y < g->h - 1 && gs_free(g, x, y + 1)) { int s = train_score(g, x, y + 1, aa->type, unit->player); if (s > best || best_dir == -1) { best = s; best_dir = O_DOWN; } }
            if (x > 0 && gs_free(g, x - 1, y)) { int s = train_score(g, x - 1, y, aa->type, unit->player); if (s > best || best_dir == -1) { best = s; best_dir = O_LEFT; } }
            aa->completed = 1;
            if (best_dir != -1) { OAct ua = mk_act(O_PRODUCE, best_dir, 0, 0, aa->type); if (is_unit_action_allowed(g, aa->unit, &ua)) { *out = ua; return 1; } }
            return 0;
        }
    }
    return 0;
}

/* AbstractionLayerAI.translateActions, AbstractionLayerAI.java:58-113 */
static int ai_translate(OAi *ai, OGame *g, int player, OPair *out) {
    int nd = 0;
    OPair *desires = (OPair *)malloc(sizeof(OPair) * (ai->n + 1));
    int *del = (int *)calloc(ai->n + 1, sizeof(int));
    ORu ru; ru_init(&ru);
    for (int i = 0; i < ai->n; i++) {
        OAbs *aa = &ai->a[i];
        if (!in_list(g, aa->unit)) { del[i] = 1; continue; }
        if (aa_completed(aa, g)) { del[i] = 1; continue; }
        if (find_assign(g, aa->unit) < 0) {
            OAct ua;
            if (aa_execute(ai, aa, g, &ru, &ua)) {
                desires[nd].unit = aa->unit; desires[nd].act = ua; nd++;
                ru_merge1(&ru, act_ru(&desires[nd - 1].act, g, aa->unit));
            }
        }
    }
    int k = 0;
    for (int i = 0; i < ai->n; i++) if (!del[i]) ai->a[k++] = ai->a[i];
    ai->n = k;
    free(del); ru_free(&ru);
    /* compose desires :93-101 */
    ORu r; ru_init(&r);
    gs_resource_usage(g, &r);
    int n = 0;
    for (int i = 0; i < nd; i++) {
        const ORu1 *r2 = act_ru(&desires[i].act, g, desires[i].unit);
        if (ru_consistent_with_ru1(&r, r2, g)) { out[n++] = desires[i]; ru_merge1(&r, r2); }
    }
    ru_free(&r); free(desires);
    /* PlayerAction.fillWithNones(gs, player, 10), PlayerAction.java:217-235 */
    for (int i = 0; i < g->n; i++) {
        int u = g->list[i];
        if (g->pool[u].player != player || find_assign(g, u) >= 0) continue;
        int found = 0;
        for (int j = 0; j < n; j++) if (out[j].unit == u) { found = 1; break; }
        if (!found) { out[n].unit = u; out[n].act = act_none(10); n++; }
    }
    return n;
}

/* PhysicalGameState.getAllFree (PhysicalGameState.java:512-525) + AbstractionLayerAI.findBuildingPosition :143-229 */
static int find_building_position(const OGame *g, const int *reserved, int nres, int dX, int dY) {
    int w = g->w, h = g->h;
    uint8_t *fr = (uint8_t *)malloc((size_t)w * h);
    for (int i = 0; i < w * h; i++) fr[i] = g->terrain[i] == 0;
    for (int i = 0; i < g->n; i++) { const OUnit *u = &g->pool[g->list[i]]; fr[u->x + u->y * w] = 0; }
    int result = -1;
    int maxl = h > w ? h : w;
#define TRY(X, Y) do { int pos = (X) + (Y) * w; int rsv = 0; for (int q = 0; q < nres; q++) if (reserved[q] == pos) rsv = 1; \
                       if (!rsv && fr[pos]) { result = pos; goto done; } } while (0)
    for (int l = 1; l < maxl; l++) {
        for (int side = 0; side < 4; side++) {
            int x, y;
            switch (side) {
                case 0: y = dY - l; if (y < 0) continue;
                    for (int dx = -l; dx <= l; dx++) { x = dX + dx; if (x < 0 || x >= w) continue; TRY(x, y); } break;
                case 1: x = dX + l; if (x >= w) continue;
                    for (int dy = -l; dy <= l; dy++) { y = dY + dy; if (y < 0 || y >= h) continue; TRY(x, y); } break;
                case 2: y = dY + l; if (y >= h) continue;
                    for (int dx = -l; dx <= l; dx++) { x = dX + dx; if (x < 0 || x >= w) continue; TRY(x, y); } break;
                case 3: x = dX - l; if (x < 0) continue;
                    for (int dy = -l; dy <= l; dy++) { y = dY + dy; if (y < 0 || y >= h) continue; TRY(x, y); } break;
            }
        }
    }
#undef TRY
done:
    free(fr);
    return result;
}

/* AbstractionLayerAI.buildIfNotAlreadyBuilding :231-245 */
static void build_if_not_already(OAi *ai, const OGame *g, int u, int type, int dX, int dY, int *reserved, int *nres) {
    OAbs *a = ai_get(ai, u);
    if (!(a && a->kind == AA_BUILD && a->type == type)) {
        int pos = find_building_position(g, reserved, *nres, dX, dY);
        ai_build(ai, u, type, pos % g->w, pos / g->w); /* Java % and / truncate like C for pos = -1 */
        reserved[(*nres)++] = pos;
    }
}

/* WorkerRush/LightRush.meleeUnitBehavior (WorkerRush.java:105-121, LightRush.java:141-159) */
/* WorkerRushPlusPlus.java is WorkerDefense.java whose melee units always attack (:116-143; its `resourse` flag never changes) */
static int is_defense(int kind) { return (kind >= O_AI_WORKER_DEFENSE && kind <= O_AI_RANGED_DEFENSE) || kind == O_AI_WORKER_RUSH_PP; }
static int is_po_rush(int kind) { return kind >= O_AI_PO_WORKER_RUSH && kind <= O_AI_PO_RANGED_RUSH; }
/* PartiallyObservableGameState.observable :61-71 */
static int po_observable(const OGame *g, int x, int y) {
    int observer = g->po_observer - 1;
    for (int i = 0; i < g->n; i++) {
        const OUnit *u = &g->pool[g->list[i]];
        if (u->player == observer) {
            int d = (u->x - x) * (u->x - x) + (u->y - y) * (u->y - y), sr = g->utt->f[u->type][OF_SIGHT];
            if (d <= sr * sr) return 1;
        }
    }
    return 0;
}

/* meleeUnitBehavior: the rushes attack the closest enemy (WorkerRush.java:105-121, LightRush.java:141-159); the defenses do
 * so only while the enemy or the own base (the LAST own base of the unit list; 0 without one) is closer than height/2, and
 * otherwise put an Attack with a null target, which translateActions deletes as completed
 * (WorkerDefense.java:117-146, LightDefense.java:142-165) */
static void melee_behavior(OAi *ai, const OGame *g, int u, int player) {
    int closest = -1, cd = 0, mybase = 0;
    const OUnit *me = &g->pool[u];
    for (int i = 0; i < g->n; i++) {
        const OUnit *o = &g->pool[g->list[i]];
        if (o->player >= 0 && o->player != player) {
            int d = abs(o->x - me->x) + abs(o->y - me->y);
            if (closest < 0 || d < cd) { closest = g->list[i]; cd = d; }
        } else if (o->player == player && o->type == 1 /* baseType */) {
            mybase = abs(o->x - me->x) + abs(o->y - me->y);
        }
    }
    if (is_defense(ai->kind)) {
        if (closest >= 0 && (ai->kind == O_AI_WORKER_RUSH_PP || cd < g->h / 2 || mybase < g->h / 2)) ai_attack(ai, u, closest);
        else ai_attack(ai, u, -1);
        return;
    }
    if (closest >= 0) { ai_attack(ai, u, closest); return; }
    if (ai->po_rush && g->po_observer) {
        /* PO*Rush.meleeUnitBehavior, partialobservability/POLightRush.java:56-77: explore the nearest non-observable cell */
        int cx = 0, cy = 0, best = -1;
        for (int i = 0; i < g->h; i++)
            for (int j = 0; j < g->w; j++)
                if (!po_observable(g, j, i)) {
                    int d = (me->x - j) * (me->x - j) + (me->y - i) * (me->y - i);
                    if (best == -1 || d < best) { cx = j; cy = i; best = d; }
                }
        if (best != -1) ai_move(ai, u, cx, cy);
    }
}

/* harvest part shared by WorkerRush.workersBehavior :148-199 and LightRush.workersBehavior :203-252; returns 1 if still free */
static int harvest_behavior(OAi *ai, const OGame *g, int u, int player) {
    const OUnit *me = &g->pool[u];
    int cbase = -1, cres = -1, cd = 0;
    for (int i = 0; i < g->n; i++) {
        const OUnit *o = &g->pool[g->list[i]];
        if (g->utt->flags[o->type] & OFL_RESOURCE) { int d = abs(o->x - me->x) + abs(o->y - me->y); if (cres < 0 || d < cd) { cres = g->list[i]; cd = d; } }
    }
    cd = 0;
    for (int i = 0; i < g->n; i++) {
        const OUnit *o = &g->pool[g->list[i]];
        if ((g->utt->flags[o->type] & OFL_STOCKPILE) && o->player == player) { int d = abs(o->x - me->x) + abs(o->y - me->y); if (cbase < 0 || d < cd) { cbase = g->list[i]; cd = d; } }
    }
    int still_free = 1;
    OAbs *aa = ai_get(ai, u);
    if (is_defense(ai->kind) || ai->kind == O_AI_CRUSH_V1 || ai->kind == O_AI_CRUSH_V2 || ai->kind == O_AI_EMR_DETERMINISTICO) { /* WorkerDefense.java:197-209, LightDefense.java:236-244, CRush_V1.java:291-320: no carrying-resources special case */
        if (cres >= 0 && cbase >= 0) {
            if (aa && aa->kind == AA_HARVEST) { if (aa->target != cres || aa->base != cbase) ai_harvest(ai, u, cres, cbase); }
            else ai_harvest(ai, u, cres, cbase);
        }
        return 0; /* the defenses never send a worker that cannot harvest to attack */
    }
    if (me->res > 0) {
        if (cbase >= 0) {
            if (aa && aa->kind == AA_HARVEST) { if (aa->base != cbase) ai_harvest(ai, u, -1, cbase); }
            else ai_harvest(ai, u, -1, cbase);
            still_free = 0;
        }
    } else {
        if (cres >= 0 && cbase >= 0) {
            if (aa && aa->kind == AA_HARVEST) { if (aa->target != cres || aa->base != cbase) ai_harvest(ai, u, cres, cbase); }
            else ai_harvest(ai, u, cres, cbase);
            still_free = 0;
        }
    }
    return still_free;
}


static int ai_get_action_k(OAi *ai, OGame *g, int player, OPair *out);
static int ai_get_action_crush(OAi *ai, OGame *g, int player, OPair *out);
static int ai_get_action_emr(OAi *ai, OGame *g, int player, OPair *out);
static int ai_get_action(OAi *ai, OGame *g, int player, OPair *out) {
    /* PO{Worker,Light,Heavy,Ranged}Rush extend their rush and override meleeUnitBehavior only (melee_behavior looks at the
     * original kind through ai->po_rush) */
    if (ai->kind == O_AI_EMR_DETERMINISTICO) return ai_get_action_emr(ai, g, player, out);
    if (ai->kind == O_AI_CRUSH_V1 || ai->kind == O_AI_CRUSH_V2) return ai_get_action_crush(ai, g, player, out);
    if (!is_po_rush(ai->kind)) return ai_get_action_k(ai, g, player, out);
    int k = ai->kind;
    ai->po_rush = 1; ai->kind = k - O_AI_PO_WORKER_RUSH + O_AI_WORKER_RUSH;
    int n = ai_get_action_k(ai, g, player, out);
    ai->kind = k; ai->po_rush = 0;
    return n;
}
static int ai_get_action_k(OAi *ai, OGame *g, int player, OPair *out) {
    const OUtt *t = g->utt;
    int BASE = type_by_role_base(), BARRACKS = type_by_role_barracks(), WORKER = type_by_role_worker();
    /* the combat unit the barracks train: Light (LightRush.java:57), Heavy (HeavyRush.java:55), Ranged (RangedRush.java:52) */
    /* the defenses: WorkerDefense.java = WorkerRush's skeleton, {Light,Heavy,Ranged}Defense.java = LightRush's, with the melee and
     * harvest rules swapped (melee_behavior / harvest_behavior) */
    int LIGHT = (ai->kind == O_AI_HEAVY_RUSH || ai->kind == O_AI_HEAVY_DEFENSE) ? 5 : ((ai->kind == O_AI_RANGED_RUSH || ai->kind == O_AI_RANGED_DEFENSE) ? 6 : type_by_role_light());
    int barracks_rush = ai->kind != O_AI_WORKER_RUSH && ai->kind != O_AI_WORKER_DEFENSE && ai->kind != O_AI_WORKER_RUSH_PP;
    int pres = g->res[player];
    /* bases: WorkerRush.java:70-76,100-102 ; LightRush.java:83-89,123-133 */
    for (int i = 0; i < g->n; i++) {
        int u = g->list[i]; const OUnit *un = &g->pool[u];
        if (un->type == BASE && un->player == player && find_assign(g, u) < 0) {
            if (!barracks_rush) {
                if (pres >= t->f[WORKER][OF_COST]) ai_train(ai, u, WORKER);
            } else {
                int nworkers = 0;
                for (int j = 0; j < g->n; j++) { const OUnit *o = &g->pool[g->list[j]]; if (o->type == WORKER && o->player == player) nworkers++; }
                if (nworkers < 1 && pres >= t->f[WORKER][OF_COST]) ai_train(ai, u, WORKER);
            }
        }
    }
    /* barracks: LightRush.java:92-98,135-139 */
    if (barracks_rush) {
        for (int i = 0; i < g->n; i++) {
            int u = g->list[i]; const OUnit *un = &g->pool[u];
            if (un->type == BARRACKS && un->player == player && find_assign(g, u) < 0) {
                if (pres >= t->f[LIGHT][OF_COST]) ai_train(ai, u, LIGHT);
            }
        }
    }
    /* melee units */
    for (int i = 0; i < g->n; i++) {
        int u = g->list[i]; const OUnit *un = &g->pool[u];
        if ((t->flags[un->type] & OFL_ATTACK) && !(t->flags[un->type] & OFL_HARVEST) && un->player == player && find_assign(g, u) < 0)
            melee_behavior(ai, g, u, player);
    }
    /* workers (all own harvesters, busy ones included) */
    int *workers = (int *)malloc(sizeof(int) * (g->n + 1)); int nw = 0;
    for (int i = 0; i < g->n; i++) { int u = g->list[i]; const OUnit *un = &g->pool[u]; if ((t->flags[un->type] & OFL_HARVEST) && un->player == player) workers[nw++] = u; }
    if (nw > 0) {
        int nbases = 0, nbarracks = 0, resourcesUsed = 0;
        for (int i = 0; i < g->n; i++) {
            const OUnit *o = &g->pool[g->list[i]];
            if (o->type == BASE && o->player == player) nbases++;
            if (o->type == BARRACKS && o->player == player) nbarracks++;
        }
        int reserved[8]; int nres = 0;
        int fw = 0; /* freeWorkers = workers[fw..nw) */
        if (nbases == 0 && fw < nw) {
            if (pres >= t->f[BASE][OF_COST] + resourcesUsed) {
                int u = workers[fw++];
                build_if_not_already(ai, g, u, BASE, g->pool[u].x, g->pool[u].y, reserved, &nres);
                resourcesUsed += t->f[BASE][OF_COST];
            }
        }
        if (barracks_rush) {
            if (nbarracks == 0) {
                if (pres >= t->f[BARRACKS][OF_COST] + resourcesUsed && fw < nw) {
                    int u = workers[fw++];
                    build_if_not_already(ai, g, u, BARRACKS, g->pool[u].x, g->pool[u].y, reserved, &nres);
                    resourcesUsed += t->f[BARRACKS][OF_COST];
                }
            }
            int *still = (int *)malloc(sizeof(int) * (nw + 1)); int ns = 0;
            for (int i = fw; i < nw; i++) if (harvest_behavior(ai, g, workers[i], player)) still[ns++] = workers[i];
            for (int i = 0; i < ns; i++) melee_behavior(ai, g, still[i], player);
            free(still);
        } else {
            /* WorkerRush.java:146-202: one harvest worker, the rest attack; a still-free harvester goes to the END of the list */
            int *fl = (int *)malloc(sizeof(int) * (nw + 1)); int nf = 0;
            int hw = -1;
            if (fw < nw) hw = workers[fw++];
            for (int i = fw; i < nw; i++) fl[nf++] = workers[i];
            if (hw >= 0 && harvest_behavior(ai, g, hw, player)) fl[nf++] = hw;
            for (int i = 0; i < nf; i++) melee_behavior(ai, g, fl[i], player);
            free(fl);
        }
    }
    free(workers);
    return ai_translate(ai, g, player, out);
}


/* ------------------------------------------------------------------------------------------------
 * CRush_V1 (ai/abstraction/cRush/CRush_V1.java:68-421): on maps of at most 144 cells a worker rush that keeps one harvester
 * per base; on larger maps nbases + 1 harvesters, one barracks, ranged units that step back from slower enemies
 * (RangedAttack.java), and every other worker fights.
 * ---------------------------------------------------------------------------------------------- */
static void crush_ranged_behavior(OAi *ai, const OGame *g, int u, int player) { /* :194-220: ONE running distance for both searches */
    const OUnit *me = &g->pool[u];
    int enemy = -1, racks = -1, cd = 0;
    for (int i = 0; i < g->n; i++) {
        const OUnit *o = &g->pool[g->list[i]];
        if (o->player >= 0 && o->player != player) {
            int d = abs(o->x - me->x) + abs(o->y - me->y);
            if (enemy < 0 || d < cd) { enemy = g->list[i]; cd = d; }
        }
        if (o->type == type_by_role_barracks() && o->player == player) {
            int d = abs(o->x - me->x) + abs(o->y - me->y);
            if (racks < 0 || d < cd) { racks = g->list[i]; cd = d; }
        }
    }
    if (enemy >= 0) ai_ranged_attack(ai, u, enemy, racks);
}

/* CRush_V2.meleeUnitBehavior :174-219 / rangedUnitBehavior :221-264: closest enemy, own barracks, own base and enemy base in one
 * pass with one shared running distance (an enemy base is tested twice: as an enemy, then as a base) */
static void crush2_combat_behavior(OAi *ai, const OGame *g, int u, int player, int ranged, int rush) {
    const OUnit *me = &g->pool[u];
    int enemy = -1, racks = -1, base = -1, ebase = -1, cd = 0;
    for (int i = 0; i < g->n; i++) {
        const OUnit *o = &g->pool[g->list[i]];
        int d = abs(o->x - me->x) + abs(o->y - me->y);
        if (o->player >= 0 && o->player != player && (enemy < 0 || d < cd)) { enemy = g->list[i]; cd = d; }
        if (o->type == type_by_role_barracks() && o->player == player && (racks < 0 || d < cd)) { racks = g->list[i]; cd = d; }
        if (o->type == type_by_role_base() && o->player == player && (base < 0 || d < cd)) { base = g->list[i]; cd = d; }
        if (o->type == type_by_role_base() && o->player != player && (ebase < 0 || d < cd)) { ebase = g->list[i]; cd = d; }
    }
    if (enemy < 0) return;
    if (!ranged && (g->time < 400 || rush)) ai_attack(ai, u, enemy);
    else ai_tactic(ai, u, enemy, base, ebase);
}

/* CRush_V2.workersBehavior :335-383: the harvesters leave a resource alone that lies closer to the enemy base than to their own */
static void crush2_harvest(OAi *ai, const OGame *g, int u, int player) {
    const OUnit *me = &g->pool[u];
    int cbase = -1, cres = -1, ebase = -1, cd = 0;
    for (int i = 0; i < g->n; i++) {
        const OUnit *o = &g->pool[g->list[i]];
        int d = abs(o->x - me->x) + abs(o->y - me->y);
        if ((g->utt->flags[o->type] & OFL_RESOURCE) && (cres < 0 || d < cd)) { cres = g->list[i]; cd = d; }
        if (o->type == type_by_role_base() && o->player != player && (ebase < 0 || d < cd)) { ebase = g->list[i]; cd = d; }
    }
    cd = 0;
    for (int i = 0; i < g->n; i++) {
        const OUnit *o = &g->pool[g->list[i]];
        if ((g->utt->flags[o->type] & OFL_STOCKPILE) && o->player == player) { int d = abs(o->x - me->x) + abs(o->y - me->y); if (cbase < 0 || d < cd) { cbase = g->list[i]; cd = d; } }
    }
    if (cres < 0) return;
    /* distance(a, b) is 0.0 when either is null (:482-485) */
    int de = ebase >= 0 ? d2_units(&g->pool[cres], &g->pool[ebase]) : 0, db = cbase >= 0 ? d2_units(&g->pool[cres], &g->pool[cbase]) : 0;
    if (de < db) return;
    if (cbase < 0) return;
    OAbs *aa = ai_get(ai, u);
    if (aa && aa->kind == AA_HARVEST) { if (aa->target != cres || aa->base != cbase) ai_harvest(ai, u, cres, cbase); }
    else ai_harvest(ai, u, cres, cbase);
}

/* workersBehavior :222-321 (rush == 0) and rushWorkersBehavior :329-416 (rush == 1) */
static void crush_workers(OAi *ai, const OGame *g, int player, const int *workers, int nw, int rush) {
    const OUtt *t = g->utt;
    int BASE = type_by_role_base(), BARRACKS = type_by_role_barracks(), WORKER = type_by_role_worker();
    int nbases = 0, nbarracks = 0, nworkers = 0, pres = g->res[player];
    ai->resources_used = 0;
    for (int i = 0; i < g->n; i++) {
        const OUnit *o = &g->pool[g->list[i]];
        if (o->player != player) continue;
        if (o->type == BASE) nbases++;
        if (o->type == BARRACKS) nbarracks++;
        if (o->type == WORKER) nworkers++;
    }
    /* freeWorkers = workers[f0, f1), battleWorkers = workers[b0, nw) */
    int f0 = 0, f1, b0;
    int keep = rush ? nbases : nbases + 1;
    if (rush && pres == 0) { f1 = 0; b0 = 0; }
    else if (nw > keep) { f1 = keep; b0 = keep; }
    else { f1 = nw; b0 = nw; }
    if (nw == 0) return; /* `workers.isEmpty()`: the list only loses elements when more than `keep` remain */
    int reserved[8], nres = 0;
    if (nbases == 0 && f0 < f1) {
        if (pres >= t->f[BASE][OF_COST]) { int u = workers[f0++]; build_if_not_already(ai, g, u, BASE, g->pool[u].x, g->pool[u].y, reserved, &nres); }
    }
    if (!rush) {
        if (nbarracks == 0 && f0 < f1 && nworkers > 1 && pres >= t->f[BARRACKS][OF_COST]) {
            int u = workers[f0++];
            build_if_not_already(ai, g, u, BARRACKS, g->pool[u].x, g->pool[u].y, reserved, &nres);
            ai->resources_used += t->f[BARRACKS][OF_COST];
            ai->building_racks = 1;
        } else ai->resources_used = t->f[BARRACKS][OF_COST] * nbarracks;
        if (nbarracks > 1) ai->building_racks = 1;
    }
    const int v2 = ai->kind == O_AI_CRUSH_V2;
    for (int i = b0; i < nw; i++) { if (v2) crush2_combat_behavior(ai, g, workers[i], player, 0, rush); else melee_behavior(ai, g, workers[i], player); }
    for (int i = f0; i < f1; i++) { if (v2 && !rush) crush2_harvest(ai, g, workers[i], player); else harvest_behavior(ai, g, workers[i], player); }
}

static int ai_get_action_crush(OAi *ai, OGame *g, int player, OPair *out) {
    const OUtt *t = g->utt;
    int BASE = type_by_role_base(), BARRACKS = type_by_role_barracks(), WORKER = type_by_role_worker(), RANGED = 6;
    int rush = g->w * g->h <= 144, pres = g->res[player];
    const int v2 = ai->kind == O_AI_CRUSH_V2;
    int *workers = (int *)malloc(sizeof(int) * (g->n + 1)); int nw = 0;
    for (int i = 0; i < g->n; i++) { int u = g->list[i]; const OUnit *un = &g->pool[u]; if ((t->flags[un->type] & OFL_HARVEST) && un->player == player) workers[nw++] = u; }
    crush_workers(ai, g, player, workers, nw, rush);
    free(workers);
    for (int i = 0; i < g->n; i++) { /* bases: rushBaseBehavior :324-326, baseBehavior :133-168 */
        int u = g->list[i]; const OUnit *un = &g->pool[u];
        if (!(un->type == BASE && un->player == player && find_assign(g, u) < 0)) continue;
        if (rush) { if (pres >= t->f[WORKER][OF_COST]) ai_train(ai, u, WORKER); continue; }
        int nbases = 0, nbarracks = 0, nworkers = 0, nranged = 0, resources = pres;
        for (int j = 0; j < g->n; j++) {
            const OUnit *o = &g->pool[g->list[j]];
            if (o->player != player) continue;
            if (o->type == RANGED) nranged++;
            if (o->type == WORKER) nworkers++;
            if (o->type == BARRACKS) nbarracks++;
            if (o->type == BASE) nbases++;
        }
        if ((nworkers < nbases + 1 && pres >= t->f[WORKER][OF_COST]) || (v2 && nranged > 6)) ai_train(ai, u, WORKER); /* CRush_V2.java:154 */
        if (ai->resources_used != t->f[BARRACKS][OF_COST] * nbarracks) resources -= t->f[BARRACKS][OF_COST];
        if (ai->building_racks && resources >= t->f[WORKER][OF_COST] + t->f[RANGED][OF_COST]) ai_train(ai, u, WORKER);
    }
    for (int i = 0; i < g->n; i++) { /* barracks :170-174 */
        int u = g->list[i]; const OUnit *un = &g->pool[u];
        if (un->type == BARRACKS && un->player == player && find_assign(g, u) < 0 && pres >= t->f[RANGED][OF_COST]) ai_train(ai, u, RANGED);
    }
    for (int i = 0; i < g->n; i++) { /* melee and ranged units :117-128 */
        int u = g->list[i]; const OUnit *un = &g->pool[u];
        if ((t->flags[un->type] & OFL_ATTACK) && !(t->flags[un->type] & OFL_HARVEST) && un->player == player && find_assign(g, u) < 0) {
            if (v2) crush2_combat_behavior(ai, g, u, player, un->type == RANGED, rush);
            else if (un->type == RANGED) crush_ranged_behavior(ai, g, u, player); else melee_behavior(ai, g, u, player);
        }
    }
    return ai_translate(ai, g, player, out);
}

/* ------------------------------------------------------------------------------------------------
 * EMRDeterministico (ai/abstraction/EMRDeterministico.java:74-358): workers until 4 (6 per base once there is a barracks),
 * barracks training Light, Ranged, Heavy in turn, further barracks and bases as resources allow, every combat unit attacks
 * the closest enemy.  Deterministic (its Random is never used); the one subtle point is otherResourcePoint (:287-311), which
 * returns the resources in HashSet order.
 * ---------------------------------------------------------------------------------------------- */
/* first element of a java.util.HashSet<Unit> filled in list order: Unit.hashCode() = (int) ID (Unit.java:548-550), HashMap
 * spreads h ^ (h >>> 16) over a power-of-two table (16 slots, doubled while size > 0.75 * slots), iteration walks the slots
 * upwards and each slot in insertion order */
static int hashset_first(const OGame *g, const int *units, int n) {
    int cap = 16;
    while (n * 4 > cap * 3) cap *= 2;
    int best = -1; uint32_t best_slot = 0;
    for (int i = 0; i < n; i++) {
        uint32_t h = (uint32_t)g->pool[units[i]].id;
        h ^= h >> 16;
        uint32_t slot = h & (uint32_t)(cap - 1);
        if (best < 0 || slot < best_slot) { best = units[i]; best_slot = slot; }
    }
    return best;
}

static int ai_get_action_emr(OAi *ai, OGame *g, int player, OPair *out) {
    const OUtt *t = g->utt;
    const int BASE = type_by_role_base(), BARRACKS = type_by_role_barracks(), WORKER = type_by_role_worker(), LIGHT = type_by_role_light(), HEAVY = 5, RANGED = 6;
    const int pres = g->res[player];
    int nworkers = 0, nbases = 0, nbarracks = 0, nlight = 0, nranged = 0, nheavy = 0;
    for (int i = 0; i < g->n; i++) {
        const OUnit *o = &g->pool[g->list[i]];
        if (o->player != player) continue;
        if (o->type == WORKER) nworkers++;
        if (o->type == BASE) nbases++;
        if (o->type == BARRACKS) nbarracks++;
        if (o->type == LIGHT) nlight++;
        if (o->type == RANGED) nranged++;
        if (o->type == HEAVY) nheavy++;
    }
    for (int i = 0; i < g->n; i++) { /* bases :133-162 */
        int u = g->list[i]; const OUnit *un = &g->pool[u];
        if (!(un->type == BASE && un->player == player && find_assign(g, u) < 0)) continue;
        int limit = nbarracks == 0 ? 4 : 6 * nbases;
        if (nworkers < limit && pres >= t->f[WORKER][OF_COST]) ai_train(ai, u, WORKER);
    }
    for (int i = 0; i < g->n; i++) { /* barracks :164-194 */
        int u = g->list[i]; const OUnit *un = &g->pool[u];
        if (!(un->type == BARRACKS && un->player == player && find_assign(g, u) < 0)) continue;
        int sum = nlight + nheavy + nranged;
        if (sum % 3 == 0 && pres >= t->f[LIGHT][OF_COST]) ai_train(ai, u, LIGHT);
        else if (sum % 3 == 1 && pres >= t->f[RANGED][OF_COST]) ai_train(ai, u, RANGED);
        else if (sum % 3 == 2 && pres >= t->f[HEAVY][OF_COST]) ai_train(ai, u, HEAVY);
    }
    /* workers :213-284: every own Worker that can harvest, busy or not */
    int *workers = (int *)malloc(sizeof(int) * (g->n + 1)); int nw = 0;
    for (int i = 0; i < g->n; i++) { int u = g->list[i]; const OUnit *un = &g->pool[u]; if ((t->flags[un->type] & OFL_HARVEST) && un->player == player && un->type == WORKER) workers[nw++] = u; }
    if (nw > 0) {
        int narmy = nlight + nranged + nheavy, used = 0, fw = 0, reserved[8], nres = 0;
        if (nbases == 0 && fw < nw && pres >= t->f[BASE][OF_COST] + used) {
            int u = workers[fw++]; build_if_not_already(ai, g, u, BASE, g->pool[u].x, g->pool[u].y, reserved, &nres); used += t->f[BASE][OF_COST];
        }
        if (nbarracks == 0 && fw < nw) {
            if (pres >= t->f[BARRACKS][OF_COST] + used) { int u = workers[fw++]; build_if_not_already(ai, g, u, BARRACKS, g->pool[u].x, g->pool[u].y, reserved, &nres); used += t->f[BARRACKS][OF_COST]; }
        } else if (nbarracks > 0 && fw < nw && narmy > 2) {
            if (pres >= t->f[BARRACKS][OF_COST] + used) { int u = workers[fw++]; build_if_not_already(ai, g, u, BARRACKS, g->pool[u].x, g->pool[u].y, reserved, &nres); used += t->f[BARRACKS][OF_COST]; }
        }
        if (nbarracks != 0) {
            /* otherResourcePoint :287-311: the resources farther than 10 (in x or in y) from every own base */
            int *other = (int *)malloc(sizeof(int) * (g->n + 1)); int no = 0;
            for (int i = 0; i < g->n; i++) {
                const OUnit *r = &g->pool[g->list[i]];
                if (!(t->flags[r->type] & OFL_RESOURCE)) continue;
                int mine = 0;
                for (int j = 0; j < g->n; j++) {
                    const OUnit *b = &g->pool[g->list[j]];
                    if (b->type == BASE && b->player == player && abs(r->x - b->x) <= 10 && abs(r->y - b->y) <= 10) mine = 1;
                }
                if (!mine) other[no++] = g->list[i];
            }
            if (no > 0 && fw < nw && pres >= t->f[BASE][OF_COST] + used) {
                const OUnit *r = &g->pool[hashset_first(g, other, no)];
                int u = workers[fw++];
                build_if_not_already(ai, g, u, BASE, r->x + 1, r->y + 1, reserved, &nres);
                used += t->f[BASE][OF_COST];
            }
            free(other);
        }
        for (int i = fw; i < nw; i++) harvest_behavior(ai, g, workers[i], player); /* harvestWorkers :325-358 */
    }
    free(workers);
    for (int i = 0; i < g->n; i++) { /* combat units :196-211 */
        int u = g->list[i]; const OUnit *un = &g->pool[u];
        if ((t->flags[un->type] & OFL_ATTACK) && !(t->flags[un->type] & OFL_HARVEST) && un->player == player && find_assign(g, u) < 0) melee_behavior(ai, g, u, player);
    }
    return ai_translate(ai, g, player, out);
}

int o_ai_get_action(OAi *ai, OGame *g, int player, int32_t *unit_idx, OActionV *acts) {
    OPair *pa = (OPair *)malloc(sizeof(OPair) * (g->n + 1));
    int n = ai_get_action(ai, g, player, pa);
    pairs_out(g, n, pa, unit_idx, acts);
    free(pa);
    return n;
}

/* ------------------------------------------------------------------------------------------------
 * Vector actions: PlayerAction.fromVectorAction (PlayerAction.java:384-417), UnitAction.fromVectorAction
 * (UnitAction.java:675-709), JNIAI.getAction fill (ai/jni/JNIAI.java:51-55)
 * ---------------------------------------------------------------------------------------------- */
int o_from_vector_action(OGame *g, int player, int n, const int32_t *vec, int fill_none_duration, int32_t *unit_idx, OActionV *acts) {
    ORu par; ru_init(&par);
    gs_resource_usage(g, &par);
    int R = o_utt_max_attack_range(g->utt) * 2 + 1, c = R / 2;
    OPair *pa = (OPair *)malloc(sizeof(OPair) * (g->n + n + 1));
    int m = 0;
    for (int k = 0; k < n; k++) {
        const int32_t *a = vec + k * 8;
        int u = unit_at(g, a[0] % g->w, a[0] / g->w);
        if (u >= 0 && g->pool[u].player == player && find_assign(g, u) < 0) {
            OAct ua = mk_act(a[1], -1, 0, 0, -1);
            switch (a[1]) {
                case O_NONE: break;
                case O_MOVE: ua.param = a[2]; break;
                case O_HARVEST: ua.param = a[3]; break;
                case O_RETURN: ua.param = a[4]; break;
                case O_PRODUCE: ua.param = a[5]; ua.utype = a[6]; break;
                case O_ATTACK: ua.x = g->pool[u].x + (a[7] % R - c); ua.y = g->pool[u].y + (a[7] / R - c); break;
            }
            if (ua.type == O_PRODUCE && (ua.utype < 0 || ua.utype >= g->utt->n)) { g->errors |= OE_BAD_ACTION; continue; }
            if (ru1_consistent_with_ru(act_ru(&ua, g, u), &par, g)) {
                ru_merge1(&par, act_ru(&ua, g, u));
                pa[m].unit = u; pa[m].act = ua; m++;
            }
        }
    }
    ru_free(&par);
    if (fill_none_duration != -9999) {
        for (int i = 0; i < g->n; i++) {
            int u = g->list[i];
            if (g->pool[u].player != player || find_assign(g, u) >= 0) continue;
            int found = 0;
            for (int j = 0; j < m; j++) if (pa[j].unit == u) { found = 1; break; }
            if (!found) { pa[m].unit = u; pa[m].act = act_none(fill_none_duration); m++; }
        }
    }
    pairs_out(g, m, pa, unit_idx, acts);
    free(pa);
    return m;
}

/* ------------------------------------------------------------------------------------------------
 * Observations (GameState.java:922-968; PartiallyObservableGameState.java:35-179), masks
 * (UnitAction.java:711-751, JNIGridnetClient.java:210-223)
 * ---------------------------------------------------------------------------------------------- */
static void observe_common(const OGame *g, int player, int32_t *out) {
    int w = g->w, h = g->h, n = w * h;
    for (int i = 0; i < g->n; i++) {
        const OUnit *u = &g->pool[g->list[i]];
        int c = u->y * w + u->x;
        out[0 * n + c] = u->hp;
        out[1 * n + c] = u->res;
        if (u->player >= 0) out[2 * n + c] = ((u->player + player) % 2) + 1;
        out[3 * n + c] = u->type + 1;
        int a = find_assign(g, g->list[i]);
        if (a >= 0) out[4 * n + c] = g->asg[a].act.type;
    }
    for (int i = 0; i < n; i++) out[5 * n + i] = g->terrain[i];
    (void)h;
}
void o_observe(const OGame *g, int player, int32_t *out) {
    memset(out, 0, sizeof(int32_t) * 6 * g->w * g->h);
    observe_common(g, player, out);
}
void o_observe_po(const OGame *g, int player, int32_t *out) {
    int w = g->w, h = g->h, n = w * h;
    memset(out, 0, sizeof(int32_t) * 8 * n);
    observe_common(g, player, out);
    for (int i = 0; i < g->n; i++) { /* calculateVisibility :156-179 */
        const OUnit *u = &g->pool[g->list[i]];
        if (u->player < 0) continue;
        int32_t *plane = out + (u->player == player ? 6 : 7) * n;
        int sr = g->utt->f[u->type][OF_SIGHT];
        for (int dy = -sr; dy <= sr; dy++) for (int dx = -sr; dx <= sr; dx++) {
            int x = u->x + dx, y = u->y + dy;
            if (x >= 0 && x < w && y >= 0 && y < h && dx * dx + dy * dy <= sr * sr) plane[y * w + x] = 1;
        }
    }
}

void o_masks(const OGame *g, int player, int32_t *out) {
    int R = o_utt_max_attack_range(g->utt) * 2 + 1, c = R / 2, nt = g->utt->n;
    int K = 1 + 6 + 16 + nt + R * R;
    memset(out, 0, sizeof(int32_t) * (size_t)g->w * g->h * K);
    OAct *l = (OAct *)malloc(sizeof(OAct) * MAX_UA);
    for (int i = 0; i < g->n; i++) {
        int u = g->list[i]; const OUnit *un = &g->pool[u];
        if (un->player != player || find_assign(g, u) >= 0) continue;
        int32_t *m = out + ((size_t)un->y * g->w + un->x) * K;
        m[0] = 1;
        int na = unit_actions(g, u, 10, l, MAX_UA);
        for (int k = 0; k < na; k++) {
            const OAct *ua = &l[k];
            m[1 + ua->type] = 1;
            switch (ua->type) {
                case O_MOVE: m[1 + 6 + ua->param] = 1; break;
                case O_HARVEST: m[1 + 6 + 4 + ua->param] = 1; break;
                case O_RETURN: m[1 + 6 + 8 + ua->param] = 1; break;
                case O_PRODUCE: m[1 + 6 + 12 + ua->param] = 1; m[1 + 6 + 16 + ua->utype] = 1; break;
                case O_ATTACK: m[1 + 6 + 16 + nt + (c + (ua->y - un->y)) * R + (c + (ua->x - un->x))] = 1; break;
            }
        }
    }
    free(l);
}

/* PartiallyObservableGameState ctor :35-54 and observable :61-71 */
OGame *o_po_view(const OGame *s, int observer) {
    OGame *g = o_game_clone(s);
    int *del = (int *)malloc(sizeof(int) * (g->n + 1)); int nd = 0;
    for (int i = 0; i < g->n; i++) {
        const OUnit *u = &g->pool[g->list[i]];
        if (u->player == observer) continue;
        int vis = 0;
        for (int j = 0; j < g->n && !vis; j++) {
            const OUnit *o = &g->pool[g->list[j]];
            if (o->player == observer) {
                int d = (o->x - u->x) * (o->x - u->x) + (o->y - u->y) * (o->y - u->y);
                int sr = g->utt->f[o->type][OF_SIGHT];
                if (d <= sr * sr) vis = 1;
            }
        }
        if (!vis) del[nd++] = g->list[i];
    }
    for (int i = 0; i < nd; i++) remove_unit(g, del[i]);
    free(del);
    g->po_observer = observer + 1;
    return g;
}

/* SimpleSqrtEvaluationFunction3.java:24-44 ; SimpleEvaluationFunction.java:21-36 */
static float base_score(const OGame *g, int fn, int player) {
    float score = g->res[player] * 20.0f;
    int any = 0;
    for (int i = 0; i < g->n; i++) {
        const OUnit *u = &g->pool[g->list[i]];
        if (u->player != player) continue;
        any = 1;
        score += u->res * 10.0f;
        int cost = g->utt->f[u->type][OF_COST], mhp = g->utt->f[u->type][OF_HP];
        if (fn == 0) {
            /* score += 40f * cost * Math.sqrt(hp / maxhp)  -- int division, float*int -> float, * double -> double, += narrows */
            double v = (double)(40.0f * (float)cost) * sqrt((double)(u->hp / mhp));
            score = (float)((double)score + v);
        } else {
            score += (40.0f * (float)(cost * u->hp)) / (float)mhp;
        }
    }
    if (fn == 0 && !any) return 0;
    return score;
}
float o_evaluate(const OGame *g, int fn, int maxplayer, int minplayer) {
    float s1 = base_score(g, fn, maxplayer), s2 = base_score(g, fn, minplayer);
    if (fn == 0) {
        if (s1 + s2 == 0) return 0.5f;
        return (2 * s1 / (s1 + s2)) - 1;
    }
    return s1 - s2;
}

/* ------------------------------------------------------------------------------------------------
 * Loops: Game.start (rts/Game.java:126-140) and NaiveMCTS.simulate (NaiveMCTS.java:297-308)
 * ---------------------------------------------------------------------------------------------- */
static int policy(OGame *g, int kind, OAi *ai, int player, OPair *out) {
    switch (kind) {
        case O_AI_RANDOM_BIASED: return rb_get_action(g, player, out);
        case O_AI_WORKER_RUSH:
        case O_AI_LIGHT_RUSH:
        case O_AI_HEAVY_RUSH:
        case O_AI_RANGED_RUSH:
        case O_AI_WORKER_DEFENSE: case O_AI_LIGHT_DEFENSE: case O_AI_HEAVY_DEFENSE:
        case O_AI_RANGED_DEFENSE:
        case O_AI_PO_WORKER_RUSH: case O_AI_PO_LIGHT_RUSH: case O_AI_PO_HEAVY_RUSH:
        case O_AI_PO_RANGED_RUSH:
        case O_AI_WORKER_RUSH_PP: case O_AI_CRUSH_V1: case O_AI_CRUSH_V2: case O_AI_EMR_DETERMINISTICO: return ai_get_action(ai, g, player, out);
        default: return 0; /* PassiveAI: empty PlayerAction */
    }
}

int o_run_game(OGame *g, int kind0, OAi *ai0, int kind1, OAi *ai1, int n_cycles, int max_cycles, int64_t *stats) {
    int gameover = o_game_gameover(g) && g->time > 0;
    OPair *p0 = NULL, *p1 = NULL; int cap = 0;
    for (int it = 0; it < n_cycles && !gameover && g->time < max_cycles; it++) {
        if (g->n + 8 > cap) { cap = g->n * 2 + 64; p0 = (OPair *)realloc(p0, sizeof(OPair) * cap); p1 = (OPair *)realloc(p1, sizeof(OPair) * cap); }
        int n0 = policy(g, kind0, ai0, 0, p0);
        int n1 = policy(g, kind1, ai1, 1, p1);
        gs_issue_safe(g, n0, p0);
        gs_issue_safe(g, n1, p1);
        gameover = o_game_cycle(g);
        if (stats) { stats[0] += 1; stats[1] += n0 + n1; stats[2] += g->n; }
    }
    free(p0); free(p1);
    return gameover;
}

/* Game.start with partiallyObservable = true (Game.java:129-140): each AI decides on new PartiallyObservableGameState(gs, p);
 * the lists are issued with issueSafe on the real state.  The view is a copy with the same unit handles, so the pairs and
 * the AIs' abstract actions refer to the real game; the policy RNG (a static in the reference) is carried back. */
static int policy_po(OGame *g, int kind, OAi *ai, int player, OPair *out) {
    if (kind != O_AI_RANDOM_BIASED && !(kind >= O_AI_WORKER_RUSH && kind <= O_AI_EMR_DETERMINISTICO)) return policy(g, kind, ai, player, out);
    OGame *v = o_po_view(g, player);
    int n = policy(v, kind, ai, player, out);
    g->rng_policy = v->rng_policy;
    o_game_free(v);
    return n;
}
int o_run_game_po(OGame *g, int kind0, OAi *ai0, int kind1, OAi *ai1, int n_cycles, int max_cycles) {
    int gameover = o_game_gameover(g) && g->time > 0;
    OPair *p0 = NULL, *p1 = NULL; int cap = 0;
    for (int it = 0; it < n_cycles && !gameover && g->time < max_cycles; it++) {
        if (g->n + 8 > cap) { cap = g->n * 2 + 64; p0 = (OPair *)realloc(p0, sizeof(OPair) * cap); p1 = (OPair *)realloc(p1, sizeof(OPair) * cap); }
        int n0 = policy_po(g, kind0, ai0, 0, p0);
        int n1 = policy_po(g, kind1, ai1, 1, p1);
        gs_issue_safe(g, n0, p0);
        gs_issue_safe(g, n1, p1);
        gameover = o_game_cycle(g);
    }
    free(p0); free(p1);
    return gameover;
}

int o_simulate(OGame *g, int time_limit) {
    int gameover = 0;
    OPair *pa = NULL; int cap = 0;
    do {
        if (o_game_is_complete(g)) {
            gameover = o_game_cycle(g);
        } else {
            if (g->n + 8 > cap) { cap = g->n * 2 + 64; pa = (OPair *)realloc(pa, sizeof(OPair) * cap); }
            int n = rb_get_action(g, 0, pa); gs_issue(g, n, pa);
            n = rb_get_action(g, 1, pa); gs_issue(g, n, pa);
        }
    } while (!gameover && g->time < time_limit);
    free(pa);
    return gameover;
}

/* bench helper (cpu_baseline of the observation workload): the Game.start loop body followed by getVectorObservation of
 * both players every cycle, as a vectorised-RL client does (JNIGridnetClientSelfPlay.gameStep :157-190).  Returns cycles run. */
int o_run_game_observing(OGame *g, int kind0, OAi *ai0, int kind1, OAi *ai1, int n_cycles, int max_cycles, int32_t *scratch) {
    int done = 0;
    for (int it = 0; it < n_cycles && g->time < max_cycles; it++) {
        int over = o_run_game(g, kind0, ai0, kind1, ai1, 1, max_cycles, NULL);
        o_observe(g, 0, scratch);
        o_observe(g, 1, scratch);
        done++;
        if (over) break;
    }
    return done;
}

/* ------------------------------------------------------------------------------------------------
 * PlayerActionGenerator (rts/PlayerActionGenerator.java) and GameState.getPlayerActions (GameState.java:493-524)
 * PARITY UNPINNED: the reference holds no golden data for them; the restatement follows the cited lines.
 * ---------------------------------------------------------------------------------------------- */
typedef struct { int unit; int n; OAct *l; } OChoice;
struct OPag {
    OGame *g; /* borrowed */
    ORu base_ru;
    int nc; OChoice *c;
    int64_t size, generated;
    int *sizes, *cur;
    int more;
};

/* PlayerActionGenerator(GameState, pID, noneDuration), PlayerActionGenerator.java:56-106; NULL = "created with no units that can execute actions" */
OPag *o_pag_create(OGame *g, int player, int none_duration) {
    OPag *p = (OPag *)calloc(1, sizeof(OPag));
    p->g = g; ru_init(&p->base_ru); gs_resource_usage(g, &p->base_ru);
    p->size = 1; p->more = 1;
    p->c = (OChoice *)calloc((size_t)g->n + 1, sizeof(OChoice));
    OAct *tmp = (OAct *)malloc(sizeof(OAct) * MAX_UA);
    for (int i = 0; i < g->n; i++) {
        int u = g->list[i];
        if (g->pool[u].player == player && find_assign(g, u) < 0) {
            int n = unit_actions(g, u, none_duration, tmp, MAX_UA);
            OChoice *c = &p->c[p->nc++];
            c->unit = u; c->n = n; c->l = (OAct *)malloc(sizeof(OAct) * (size_t)n); memcpy(c->l, tmp, sizeof(OAct) * (size_t)n);
            if (INT64_MAX / p->size <= (int64_t)n) p->size = INT64_MAX; else p->size *= (int64_t)n;
        }
    }
    free(tmp);
    if (p->nc == 0) { ru_free(&p->base_ru); free(p->c); free(p); return NULL; }
    p->sizes = (int *)malloc(sizeof(int) * (size_t)p->nc); p->cur = (int *)calloc((size_t)p->nc, sizeof(int));
    for (int i = 0; i < p->nc; i++) p->sizes[i] = p->c[i].n;
    return p;
}
void o_pag_free(OPag *p) {
    if (!p) return;
    for (int i = 0; i < p->nc; i++) free(p->c[i].l);
    ru_free(&p->base_ru); free(p->c); free(p->sizes); free(p->cur); free(p);
}
int64_t o_pag_size(const OPag *p) { return p->size; }
int64_t o_pag_generated(const OPag *p) { return p->generated; }
int o_pag_n_choices(const OPag *p) { return p->nc; }

static void ru_clone(ORu *dst, const ORu *src) { ru_init(dst); for (int i = 0; i < src->npos; i++) ru_add_pos(dst, src->pos[i]); dst->res[0] = src->res[0]; dst->res[1] = src->res[1]; }

/* incrementCurrentChoice, PlayerActionGenerator.java:128-140 */
static void pag_increment(OPag *p, int start) {
    for (int i = 0; i < start; i++) p->cur[i] = 0;
    p->cur[start]++;
    if (p->cur[start] >= p->sizes[start]) {
        if (start < p->nc - 1) pag_increment(p, start + 1); else p->more = 0;
    }
}
/* getNextAction, PlayerActionGenerator.java:148-195 (no cut-off time): pairs in the order they are added (last choice first);
 * returns the number of pairs, or -1 when the generator is exhausted */
int o_pag_next(OPag *p, int32_t *unit_idx, OActionV *acts) {
    OGame *g = p->g;
    while (p->more) {
        int consistent = 1, n = 0;
        ORu r; ru_clone(&r, &p->base_ru);
        int i = p->nc;
        while (i > 0) {
            i--;
            OChoice *c = &p->c[i];
            OAct *ua = &c->l[p->cur[i]];
            const ORu1 *r2 = act_ru(ua, g, c->unit);
            if (ru_consistent_with_ru1(&r, r2, g)) { ru_merge1(&r, r2); unit_idx[n] = list_index_of(g, c->unit); acts[n] = act_to_v(ua); n++; }
            else { consistent = 0; break; }
        }
        ru_free(&r);
        pag_increment(p, i);
        if (consistent) { p->generated++; return n; }
    }
    return -1;
}
/* randomizeOrder, PlayerActionGenerator.java:114-121 with the given Random in place of the static one */
void o_pag_randomize_order(OPag *p, OJRandom *r) {
    for (int i = 0; i < p->nc; i++) {
        OChoice *c = &p->c[i];
        int m = c->n;
        OAct *tmp = (OAct *)malloc(sizeof(OAct) * (size_t)m); memcpy(tmp, c->l, sizeof(OAct) * (size_t)m);
        for (int k = 0; k < c->n; k++) {
            int j = o_jr_next_int_bound(r, m);
            c->l[k] = tmp[j];
            memmove(&tmp[j], &tmp[j + 1], sizeof(OAct) * (size_t)(m - j - 1)); m--;
        }
        free(tmp);
    }
}
/* getRandom, PlayerActionGenerator.java:201-222 with the given Random in place of `new Random()` */
int o_pag_random(OPag *p, OJRandom *r, int32_t *unit_idx, OActionV *acts) {
    OGame *g = p->g;
    ORu ru; ru_clone(&ru, &p->base_ru);
    int n = 0;
    for (int i = 0; i < p->nc; i++) {
        OChoice *c = &p->c[i];
        int m = c->n;
        OAct *l = (OAct *)malloc(sizeof(OAct) * (size_t)m); memcpy(l, c->l, sizeof(OAct) * (size_t)m);
        int consistent = 0;
        do {
            int j = o_jr_next_int_bound(r, m);
            OAct ua = l[j];
            memmove(&l[j], &l[j + 1], sizeof(OAct) * (size_t)(m - j - 1)); m--;
            const ORu1 *r2 = act_ru(&ua, g, c->unit);
            if (ru_consistent_with_ru1(&ru, r2, g)) { ru_merge1(&ru, r2); unit_idx[n] = list_index_of(g, c->unit); acts[n] = act_to_v(&ua); n++; consistent = 1; }
        } while (!consistent);
        free(l);
    }
    ru_free(&ru);
    return n;
}

/* GameState.getPlayerActions, GameState.java:493-524 + PlayerAction.cartesianProduct, PlayerAction.java:180-195.
 * out: per PlayerAction {n, n x (unit list index, type, param, x, y, utype)}; returns the number of PlayerActions (writes while they fit) */
int64_t o_player_actions(OGame *g, int player, int32_t *out, int64_t max_ints) {
    typedef struct { ORu r; int n; OPair *a; } PA;
    int64_t nl = 1, cap = 16;
    PA *l = (PA *)malloc(sizeof(PA) * (size_t)cap);
    ru_init(&l[0].r); gs_resource_usage(g, &l[0].r); l[0].n = 0; l[0].a = NULL;
    OAct *tmp = (OAct *)malloc(sizeof(OAct) * MAX_UA);
    for (int i = 0; i < g->n; i++) {
        int u = g->list[i];
        if (g->pool[u].player != player || find_assign(g, u) >= 0) continue;
        int na = unit_actions(g, u, 10, tmp, MAX_UA);
        int64_t n2 = 0, cap2 = nl * 2 + 16;
        PA *l2 = (PA *)malloc(sizeof(PA) * (size_t)cap2);
        for (int64_t k = 0; k < nl; k++) {
            for (int a = 0; a < na; a++) {
                const ORu1 *r2 = act_ru(&tmp[a], g, u);
                if (!ru_consistent_with_ru1(&l[k].r, r2, g)) continue;
                if (n2 == cap2) { cap2 *= 2; l2 = (PA *)realloc(l2, sizeof(PA) * (size_t)cap2); }
                PA *q = &l2[n2++];
                ru_clone(&q->r, &l[k].r); ru_merge1(&q->r, r2);
                q->n = l[k].n + 1; q->a = (OPair *)malloc(sizeof(OPair) * (size_t)q->n);
                if (l[k].n) memcpy(q->a, l[k].a, sizeof(OPair) * (size_t)l[k].n);
                q->a[q->n - 1].unit = u; q->a[q->n - 1].act = tmp[a];
            }
        }
        for (int64_t k = 0; k < nl; k++) { ru_free(&l[k].r); free(l[k].a); }
        free(l); l = l2; nl = n2;
    }
    free(tmp);
    int64_t w = 0;
    for (int64_t k = 0; k < nl; k++) {
        if (w + 1 + 6 * (int64_t)l[k].n <= max_ints) {
            out[w++] = l[k].n;
            for (int j = 0; j < l[k].n; j++) {
                out[w++] = list_index_of(g, l[k].a[j].unit); out[w++] = l[k].a[j].act.type; out[w++] = l[k].a[j].act.param;
                out[w++] = l[k].a[j].act.x; out[w++] = l[k].a[j].act.y; out[w++] = l[k].a[j].act.utype;
            }
        }
        ru_free(&l[k].r); free(l[k].a);
    }
    free(l);
    return nl;
}

/* ------------------------------------------------------------------------------------------------
 * NaiveMCTS (ai/mcts/naivemcts/NaiveMCTS.java, NaiveMCTSNode.java, ai/mcts/MCTSNode.java), RandomBiasedAI playouts,
 * SimpleSqrtEvaluationFunction3 / SimpleEvaluationFunction.  The reference draws from two unseeded static generators (MCTSNode.r
 * and util.Sampler.generator, the latter shared with the playout policy) and cannot be replayed; here a search owns a seeded
 * MCTSNode.r stream and a seeded Sampler stream for the tree, and playout k of the search is seeded with seed * 1000003 + k.
 * PARITY UNPINNED (no golden data in the reference).
 * ---------------------------------------------------------------------------------------------- */
typedef struct ONode {
    int type;            /* 0 max, 1 min, -1 game over */
    struct ONode *parent;
    OGame *gs;
    int depth;
    double accum; int visits;
    int nch, chcap; struct ONode **children;
    int **codes;         /* per child: the code vector (one action index per choice) == the BigInteger action code */
    OPair **pas; int *pan; /* per child: the PlayerAction in the order it was built */
    OPag *gen;           /* moveGenerator; its choices are the unit action table */
    double **ate_accum; int **ate_visits;
} ONode;
struct OMcts {
    ONode *tree; OGame *start; int player;
    int lookahead, max_depth, strategy, fensa, eval_fn;
    float e_l, e_g, e_0;
    OJRandom r, sampler;
    int64_t seed; int64_t runs;
    double bound;
};

static float jr_next_float(OJRandom *r) { return o_jr_next(r, 24) / (float)(1 << 24); }

static ONode *node_new(OMcts *m, OGame *gs /* owned */, ONode *parent) {
    ONode *nd = (ONode *)calloc(1, sizeof(ONode));
    nd->parent = parent; nd->gs = gs; nd->depth = parent ? parent->depth + 1 : 0;
    int maxp = m->player, minp = 1 - m->player;
    while (o_game_winner(gs) == -1 && !o_game_gameover(gs) && !can_execute_any(gs, maxp) && !can_execute_any(gs, minp)) o_game_cycle(gs);
    if (o_game_winner(gs) != -1 || o_game_gameover(gs)) nd->type = -1;
    else if (can_execute_any(gs, maxp)) { nd->type = 0; nd->gen = o_pag_create(gs, maxp, 10); }
    else if (can_execute_any(gs, minp)) { nd->type = 1; nd->gen = o_pag_create(gs, minp, 10); }
    else nd->type = -1;
    if (nd->gen) {
        nd->ate_accum = (double **)malloc(sizeof(double *) * (size_t)nd->gen->nc); nd->ate_visits = (int **)malloc(sizeof(int *) * (size_t)nd->gen->nc);
        for (int i = 0; i < nd->gen->nc; i++) { nd->ate_accum[i] = (double *)calloc((size_t)nd->gen->c[i].n, sizeof(double)); nd->ate_visits[i] = (int *)calloc((size_t)nd->gen->c[i].n, sizeof(int)); }
    }
    return nd;
}
static void node_free(ONode *nd) {
    if (!nd) return;
    for (int i = 0; i < nd->nch; i++) { node_free(nd->children[i]); free(nd->codes[i]); free(nd->pas[i]); }
    if (nd->gen) { for (int i = 0; i < nd->gen->nc; i++) { free(nd->ate_accum[i]); free(nd->ate_visits[i]); } free(nd->ate_accum); free(nd->ate_visits); }
    o_pag_free(nd->gen); o_game_free(nd->gs);
    free(nd->children); free(nd->codes); free(nd->pas); free(nd->pan); free(nd);
}

/* Sampler.weighted(double[]), util/Sampler.java:116-137 on the search's Sampler stream */
static int sampler_weighted(OMcts *m, const double *dist, int n) {
    double total = 0, accum = 0, tmp;
    for (int i = 0; i < n; i++) total += dist[i];
    if (total == 0) return o_jr_next_int_bound(&m->sampler, n);
    tmp = o_jr_next_double(&m->sampler) * total;
    for (int i = 0; i < n; i++) { accum += dist[i]; if (accum >= tmp) return i; }
    return n - 1; /* Java: throws */
}

static ONode *select_leaf(OMcts *m, ONode *nd);

/* selectFromAlreadySampledEpsilonGreedy, NaiveMCTSNode.java:136-162 */
static ONode *select_egreedy(OMcts *m, ONode *nd) {
    if (jr_next_float(&m->r) >= m->e_g) {
        ONode *best = NULL;
        for (int i = 0; i < nd->nch; i++) {
            ONode *c = nd->children[i];
            if (nd->type == 0) { if (!best || (c->accum / c->visits) > (best->accum / best->visits)) best = c; }
            else { if (!best || (c->accum / c->visits) < (best->accum / best->visits)) best = c; }
        }
        return best;
    }
    return nd->children[o_jr_next_int_bound(&m->r, nd->nch)];
}
/* selectFromAlreadySampledUCB1, NaiveMCTSNode.java:165-188 (C = 0.05) */
static ONode *select_ucb1(OMcts *m, ONode *nd) {
    ONode *best = NULL; double bestScore = 0; const float C = 0.05f;
    for (int i = 0; i < nd->nch; i++) {
        ONode *c = nd->children[i];
        double exploitation = ((double)c->accum) / c->visits;
        double exploration = sqrt(log((double)nd->visits) / c->visits);
        if (nd->type == 0) exploitation = (m->bound + exploitation) / (2 * m->bound);
        else exploitation = (m->bound - exploitation) / (2 * m->bound);
        double tmp = C * exploitation + exploration;
        if (!best || tmp > bestScore) { best = c; bestScore = tmp; }
    }
    return best;
}

/* selectLeafUsingLocalMABs, NaiveMCTSNode.java:191-330 */
static ONode *select_local(OMcts *m, ONode *nd) {
    OPag *gen = nd->gen; OGame *gs = nd->gs;
    int nc = gen->nc;
    double **dists = (double **)malloc(sizeof(double *) * (size_t)nc);
    int *not_sampled = (int *)malloc(sizeof(int) * (size_t)nc), nns = 0;
    for (int e = 0; e < nc; e++) {
        int na = gen->c[e].n;
        double *dist = (double *)malloc(sizeof(double) * (size_t)na);
        int bestIdx = -1, visits = 0; double bestEval = 0;
        const double *acc = nd->ate_accum[e]; const int *vc = nd->ate_visits[e];
        for (int i = 0; i < na; i++) {
            if (nd->type == 0) {
                if (bestIdx == -1 || (visits != 0 && vc[i] == 0) || (visits != 0 && (acc[i] / vc[i]) > bestEval)) {
                    bestIdx = i; bestEval = vc[i] > 0 ? acc[i] / vc[i] : 0; visits = vc[i];
                }
            } else {
                if (bestIdx == -1 || (visits != 0 && vc[i] == 0) || (visits != 0 && (acc[i] / vc[i]) < bestEval)) {
                    bestIdx = i; bestEval = vc[i] > 0 ? acc[i] / vc[i] : 0; visits = vc[i];
                }
            }
            dist[i] = m->e_l / na; /* float arithmetic, widened */
        }
        if (vc[bestIdx] != 0) dist[bestIdx] = (1 - m->e_l) + (m->e_l / na);
        else if (m->fensa) { for (int j = 0; j < na; j++) if (vc[j] > 0) dist[j] = 0; }
        not_sampled[nns++] = e;
        dists[e] = dist;
    }
    ORu ru; ru_init(&ru); gs_resource_usage(gs, &ru);
    OPair *pa = (OPair *)malloc(sizeof(OPair) * (size_t)nc); int npa = 0;
    int *code_v = (int *)calloc((size_t)nc, sizeof(int));
    while (nns > 0) {
        int k = o_jr_next_int_bound(&m->r, nns);
        int i = not_sampled[k];
        memmove(&not_sampled[k], &not_sampled[k + 1], sizeof(int) * (size_t)(nns - k - 1)); nns--;
        OChoice *c = &gen->c[i];
        const double *distribution = dists[i];
        int code = sampler_weighted(m, distribution, c->n);
        OAct *ua = &c->l[code];
        const ORu1 *r2 = act_ru(ua, gs, c->unit);
        if (!ru_consistent_with_ru1(&ru, r2, gs)) {
            int nl = c->n;
            double *dl = (double *)malloc(sizeof(double) * (size_t)nl); int *outs = (int *)malloc(sizeof(int) * (size_t)nl);
            for (int j = 0; j < nl; j++) { dl[j] = distribution[j]; outs[j] = j; }
            do {
                int idx = 0; while (outs[idx] != code) idx++;
                memmove(&dl[idx], &dl[idx + 1], sizeof(double) * (size_t)(nl - idx - 1));
                memmove(&outs[idx], &outs[idx + 1], sizeof(int) * (size_t)(nl - idx - 1)); nl--;
                /* Sampler.weighted(List<Double>, List<?>), util/Sampler.java:141-161 */
                double total = 0, accum = 0, tmp;
                for (int j = 0; j < nl; j++) total += dl[j];
                if (total == 0) code = outs[o_jr_next_int_bound(&m->sampler, nl)];
                else { tmp = o_jr_next_double(&m->sampler) * total; code = outs[nl - 1]; for (int j = 0; j < nl; j++) { accum += dl[j]; if (accum >= tmp) { code = outs[j]; break; } } }
                ua = &c->l[code];
                r2 = act_ru(ua, gs, c->unit);
            } while (!ru_consistent_with_ru1(&ru, r2, gs));
            free(dl); free(outs);
        }
        ru_merge1(&ru, r2);
        pa[npa].unit = c->unit; pa[npa].act = *ua; npa++;
        code_v[i] = code;
    }
    ru_free(&ru);
    for (int e = 0; e < nc; e++) free(dists[e]);
    free(dists); free(not_sampled);
    for (int ch = 0; ch < nd->nch; ch++)
        if (memcmp(nd->codes[ch], code_v, sizeof(int) * (size_t)nc) == 0) { free(code_v); free(pa); return select_leaf(m, nd->children[ch]); }
    /* new child: gs.cloneIssue(pa2), then the node's own loop */
    OGame *gs2 = o_game_clone(gs);
    OPair *keep = (OPair *)malloc(sizeof(OPair) * (size_t)(npa ? npa : 1)); /* `actions.add(pa2)`: the PlayerAction as it was built */
    memcpy(keep, pa, sizeof(OPair) * (size_t)npa);
    gs_issue(gs2, npa, pa);
    if (nd->nch == nd->chcap) {
        nd->chcap = nd->chcap ? nd->chcap * 2 : 8;
        nd->children = (ONode **)realloc(nd->children, sizeof(ONode *) * (size_t)nd->chcap); nd->codes = (int **)realloc(nd->codes, sizeof(int *) * (size_t)nd->chcap);
        nd->pas = (OPair **)realloc(nd->pas, sizeof(OPair *) * (size_t)nd->chcap); nd->pan = (int *)realloc(nd->pan, sizeof(int) * (size_t)nd->chcap);
    }
    ONode *node = node_new(m, gs2, nd);
    nd->children[nd->nch] = node; nd->codes[nd->nch] = code_v; nd->pas[nd->nch] = keep; nd->pan[nd->nch] = npa; nd->nch++;
    free(pa);
    return node;
}

/* selectLeaf, NaiveMCTSNode.java:108-133 */
static ONode *select_leaf(OMcts *m, ONode *nd) {
    if (!nd->gen) return nd;
    if (nd->depth >= m->max_depth) return nd;
    if (nd->nch > 0 && jr_next_float(&m->r) >= m->e_0) {
        ONode *sel = m->strategy == 0 ? select_egreedy(m, nd) : select_ucb1(m, nd);
        return select_leaf(m, sel);
    }
    return select_local(m, nd);
}

/* propagateEvaluation, NaiveMCTSNode.java:341-368 */
static void propagate(ONode *nd, double evaluation, ONode *child) {
    nd->accum += evaluation; nd->visits++;
    if (child) {
        int idx = 0; while (nd->children[idx] != child) idx++;
        const int *code = nd->codes[idx];
        for (int i = 0; i < nd->gen->nc; i++) { nd->ate_accum[i][code[i]] += evaluation; nd->ate_visits[i][code[i]]++; }
    }
    if (nd->parent) propagate(nd->parent, evaluation, nd);
}

OMcts *o_mcts_create(const OGame *g, int player, int lookahead, int max_depth, float e_l, float e_g, float e_0, int strategy, int fensa, int eval_fn, int64_t seed) {
    OMcts *m = (OMcts *)calloc(1, sizeof(OMcts));
    m->player = player; m->lookahead = lookahead; m->max_depth = max_depth; m->e_l = e_l; m->e_g = e_g; m->e_0 = e_0; m->strategy = strategy; m->fensa = fensa;
    m->eval_fn = eval_fn; m->seed = seed; m->bound = 1.0;
    o_jr_seed(&m->r, seed); o_jr_seed(&m->sampler, seed ^ 0x2545F4914F6CDD1DLL);
    m->tree = node_new(m, o_game_clone(g), NULL); /* startNewComputation(player, gs.clone()): tree = new NaiveMCTSNode(..., gs, ...) */
    m->start = o_game_clone(m->tree->gs);         /* gs_to_start_from = gs: the same object, i.e. after the root's own cycle loop */
    return m;
}
void o_mcts_free(OMcts *m) { if (!m) return; node_free(m->tree); o_game_free(m->start); free(m); }

/* NaiveMCTS.iteration, NaiveMCTS.java:195-223, n times */
void o_mcts_iterate(OMcts *m, int n) {
    for (int it = 0; it < n; it++) {
        ONode *leaf = select_leaf(m, m->tree);
        OGame *gs2 = o_game_clone(leaf->gs);
        o_game_seed(gs2, m->seed * 1000003LL + m->runs);
        o_simulate(gs2, gs2->time + m->lookahead);
        int time = gs2->time - m->start->time;
        double evaluation = o_evaluate(gs2, m->eval_fn, m->player, 1 - m->player) * pow(0.99, time / 10.0);
        o_game_free(gs2);
        propagate(leaf, evaluation, NULL);
        m->runs++;
    }
}
/* root statistics: out_visits/out_accum per child (creation order); returns the number of children */
int o_mcts_root(const OMcts *m, int *root_visits, double *root_accum, int *out_visits, double *out_accum, int max_children) {
    *root_visits = m->tree->visits; *root_accum = m->tree->accum;
    for (int i = 0; i < m->tree->nch && i < max_children; i++) { out_visits[i] = m->tree->children[i]->visits; out_accum[i] = m->tree->children[i]->accum; }
    return m->tree->nch;
}
int o_mcts_n_nodes_rec(const ONode *nd) { int c = 1; for (int i = 0; i < nd->nch; i++) c += o_mcts_n_nodes_rec(nd->children[i]); return c; }
int o_mcts_n_nodes(const OMcts *m) { return o_mcts_n_nodes_rec(m->tree); }
/* getBestActionSoFar / getMostVisitedActionIdx, NaiveMCTS.java:226-262: pairs of the most visited child (first maximum); -1 = no children */
int o_mcts_best_action(const OMcts *m, int32_t *unit_idx, OActionV *acts) {
    const ONode *t = m->tree; int best = -1;
    for (int i = 0; i < t->nch; i++) if (best == -1 || t->children[i]->visits > t->children[best]->visits) best = i;
    if (best < 0) return -1;
    for (int k = 0; k < t->pan[best]; k++) { unit_idx[k] = list_index_of(t->gs, t->pas[best][k].unit); acts[k] = act_to_v(&t->pas[best][k].act); }
    return t->pan[best];
}

/* ------------------------------------------------------------------------------------------------
 * UCT (ai/mcts/uct/UCT.java:103-175, UCTNode.java:37-125) with RandomBiasedAI playouts.  As for NaiveMCTS the reference's generators
 * are unseeded statics (PlayerActionGenerator.r shuffles every node's move generator); here a search owns one seeded stream for the
 * shuffles and playout k is seeded with seed * 1000003 + k.  PARITY UNPINNED.
 * ---------------------------------------------------------------------------------------------- */
typedef struct OUNode {
    int type; struct OUNode *parent; OGame *gs; int depth;
    int has_more; OPag *gen;
    int nch, chcap; struct OUNode **children; OPair **pas; int *pan;
    float accum; int visits;
} OUNode;
struct OUct { OUNode *tree; int start_time; int player, lookahead, max_depth, eval_fn; float bound; OJRandom r; int64_t seed, runs; };

static OUNode *unode_new(OUct *m, OGame *gs, OUNode *parent) {
    OUNode *nd = (OUNode *)calloc(1, sizeof(OUNode));
    nd->parent = parent; nd->gs = gs; nd->depth = parent ? parent->depth + 1 : 0; nd->has_more = 1;
    int maxp = m->player, minp = 1 - m->player;
    while (o_game_winner(gs) == -1 && !o_game_gameover(gs) && !can_execute_any(gs, maxp) && !can_execute_any(gs, minp)) o_game_cycle(gs);
    if (o_game_winner(gs) != -1 || o_game_gameover(gs)) nd->type = -1;
    else if (can_execute_any(gs, maxp)) { nd->type = 0; nd->gen = o_pag_create(gs, maxp, 10); o_pag_randomize_order(nd->gen, &m->r); }
    else if (can_execute_any(gs, minp)) { nd->type = 1; nd->gen = o_pag_create(gs, minp, 10); o_pag_randomize_order(nd->gen, &m->r); }
    else nd->type = -1;
    return nd;
}
static void unode_free(OUNode *nd) {
    if (!nd) return;
    for (int i = 0; i < nd->nch; i++) { unode_free(nd->children[i]); free(nd->pas[i]); }
    o_pag_free(nd->gen); o_game_free(nd->gs); free(nd->children); free(nd->pas); free(nd->pan); free(nd);
}
/* childValue, UCTNode.java:112-125 (C = 0.05; evaluation_bound is a float) */
static double uct_child_value(const OUct *m, const OUNode *nd, const OUNode *c) {
    const float C = 0.05f;
    double exploitation = ((double)c->accum) / c->visits;
    double exploration = sqrt(log((double)nd->visits) / c->visits);
    if (nd->type == 0) exploitation = (m->bound + exploitation) / (2 * m->bound);
    else exploitation = (m->bound - exploitation) / (2 * m->bound);
    return C * exploitation + exploration;
}
/* UCTSelectLeaf, UCTNode.java:70-109 */
static OUNode *uct_select_leaf(OUct *m, OUNode *nd) {
    if (nd->depth >= m->max_depth) return nd;
    if (nd->has_more) {
        if (!nd->gen) return nd;
        int32_t *idx = (int32_t *)malloc(sizeof(int32_t) * (size_t)(nd->gs->n + 8));
        OActionV *acts = (OActionV *)malloc(sizeof(OActionV) * (size_t)(nd->gs->n + 8));
        int n = o_pag_next(nd->gen, idx, acts);
        if (n >= 0) {
            OPair *pa = (OPair *)malloc(sizeof(OPair) * (size_t)(n ? n : 1)), *keep = (OPair *)malloc(sizeof(OPair) * (size_t)(n ? n : 1));
            for (int k = 0; k < n; k++) { pa[k].unit = nd->gs->list[idx[k]]; pa[k].act = act_from_v(&acts[k]); keep[k] = pa[k]; }
            OGame *gs2 = o_game_clone(nd->gs);
            gs_issue(gs2, n, pa);
            free(pa); free(idx); free(acts);
            if (nd->nch == nd->chcap) {
                nd->chcap = nd->chcap ? nd->chcap * 2 : 8;
                nd->children = (OUNode **)realloc(nd->children, sizeof(OUNode *) * (size_t)nd->chcap);
                nd->pas = (OPair **)realloc(nd->pas, sizeof(OPair *) * (size_t)nd->chcap); nd->pan = (int *)realloc(nd->pan, sizeof(int) * (size_t)nd->chcap);
            }
            OUNode *node = unode_new(m, gs2, nd);
            nd->children[nd->nch] = node; nd->pas[nd->nch] = keep; nd->pan[nd->nch] = n; nd->nch++;
            return node;
        }
        free(idx); free(acts);
        nd->has_more = 0;
    }
    double best_score = 0; OUNode *best = NULL;
    for (int i = 0; i < nd->nch; i++) {
        double tmp = uct_child_value(m, nd, nd->children[i]);
        if (!best || tmp > best_score) { best = nd->children[i]; best_score = tmp; }
    }
    if (!best) return nd;
    return uct_select_leaf(m, best);
}
OUct *o_uct_create(const OGame *g, int player, int lookahead, int max_depth, int eval_fn, int64_t seed) {
    OUct *m = (OUct *)calloc(1, sizeof(OUct));
    m->player = player; m->lookahead = lookahead; m->max_depth = max_depth; m->eval_fn = eval_fn; m->seed = seed; m->bound = 1.0f;
    o_jr_seed(&m->r, seed);
    m->tree = unode_new(m, o_game_clone(g), NULL);
    m->start_time = m->tree->gs->time; /* gs_to_start_from is the root's own state object */
    return m;
}
void o_uct_free(OUct *m) { if (!m) return; unode_free(m->tree); free(m); }
/* monteCarloRun, UCT.java:140-168, n times */
void o_uct_iterate(OUct *m, int n) {
    for (int it = 0; it < n; it++) {
        OUNode *leaf = uct_select_leaf(m, m->tree);
        OGame *gs2 = o_game_clone(leaf->gs);
        o_game_seed(gs2, m->seed * 1000003LL + m->runs);
        o_simulate(gs2, gs2->time + m->lookahead);
        int time = gs2->time - m->start_time;
        double evaluation = o_evaluate(gs2, m->eval_fn, m->player, 1 - m->player) * pow(0.99, time / 10.0);
        o_game_free(gs2);
        while (leaf) { leaf->accum += evaluation; /* float += double */ leaf->visits++; leaf = leaf->parent; }
        m->runs++;
    }
}
int o_uct_root(const OUct *m, int *root_visits, float *root_accum, int *out_visits, float *out_accum, int max_children) {
    *root_visits = m->tree->visits; *root_accum = m->tree->accum;
    for (int i = 0; i < m->tree->nch && i < max_children; i++) { out_visits[i] = m->tree->children[i]->visits; out_accum[i] = m->tree->children[i]->accum; }
    return m->tree->nch;
}
static int uct_count(const OUNode *nd) { int c = 1; for (int i = 0; i < nd->nch; i++) c += uct_count(nd->children[i]); return c; }
int o_uct_n_nodes(const OUct *m) { return uct_count(m->tree); }
/* getBestActionSoFar, UCT.java:171-199: most visited child, ties by accumulated evaluation */
int o_uct_best_action(const OUct *m, int32_t *unit_idx, OActionV *acts) {
    const OUNode *t = m->tree; int best = -1;
    for (int i = 0; i < t->nch; i++) {
        const OUNode *c = t->children[i];
        if (best == -1 || c->visits > t->children[best]->visits || (c->visits == t->children[best]->visits && c->accum > t->children[best]->accum)) best = i;
    }
    if (best < 0) return -1;
    for (int k = 0; k < t->pan[best]; k++) { unit_idx[k] = list_index_of(t->gs, t->pas[best][k].unit); acts[k] = act_to_v(&t->pas[best][k].act); }
    return t->pan[best];
}
