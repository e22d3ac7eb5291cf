/*
 * mrts_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE ONLY, NOT PART OF THE PRODUCT).
 *
 * A plain-C restatement of the microRTS game rules of ConnAALL/MicroRTS, following the Java
 * sources function by function (each function in mrts_oracle.c cites the file:line it follows).
 * It exists so that the CUDA path can be checked for bit-exact parity.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product library (libmicrorts_cuda.so) never links, loads or calls anything in oracle/.
 *
 * Pinning: the rules (issueSafe/issue/cycle/execute/death/winner) and LightRush+AbstractionLayerAI+A*
 * are pinned by the reference's 280 golden traces (tests/golden/traces.pack.gz).  RandomBiasedAI,
 * WorkerRush, BFS, observation planes, masks, partial observability, simulate()/evaluation, UTT v2/v3
 * and conflict policies 2/3 have no golden data in the reference: PARITY UNPINNED for those.
 */
#ifndef MRTS_ORACLE_H
#define MRTS_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { O_NONE = 0, O_MOVE = 1, O_HARVEST = 2, O_RETURN = 3, O_PRODUCE = 4, O_ATTACK = 5 };
enum { O_UP = 0, O_RIGHT = 1, O_DOWN = 2, O_LEFT = 3 };
enum { O_AI_NONE = 0, O_AI_PASSIVE = 1, O_AI_RANDOM_BIASED = 2, O_AI_WORKER_RUSH = 3, O_AI_LIGHT_RUSH = 4,
       O_AI_HEAVY_RUSH = 5, O_AI_RANGED_RUSH = 6, /* HeavyRush.java / RangedRush.java are LightRush.java with the trained type swapped */
       O_AI_WORKER_DEFENSE = 7, O_AI_LIGHT_DEFENSE = 8, O_AI_HEAVY_DEFENSE = 9, O_AI_RANGED_DEFENSE = 10,
       O_AI_PO_WORKER_RUSH = 11, O_AI_PO_LIGHT_RUSH = 12, O_AI_PO_HEAVY_RUSH = 13, O_AI_PO_RANGED_RUSH = 14, /* ai/abstraction/partialobservability/ */
       O_AI_WORKER_RUSH_PP = 15 /* ai/abstraction/WorkerRushPlusPlus.java */,
       O_AI_CRUSH_V1 = 16, O_AI_CRUSH_V2 = 17, /* ai/abstraction/cRush/CRush_V1.java, CRush_V2.java + CRanged_Tactic.java */
       O_AI_EMR_DETERMINISTICO = 18 /* ai/abstraction/EMRDeterministico.java */ }; /* ai/abstraction/{Worker,Light,Heavy,Ranged}Defense.java */
enum { O_PF_ASTAR = 0, O_PF_BFS = 1, O_PF_GREEDY = 2 };

/* unit type fields, in the order of the UTT XML attributes */
enum { OF_COST = 0, OF_HP, OF_MINDMG, OF_MAXDMG, OF_RANGE, OF_PRODUCE_T, OF_MOVE_T, OF_ATTACK_T, OF_HARVEST_T,
       OF_RETURN_T, OF_HARVEST_AMT, OF_SIGHT, OF_NFIELDS };
enum { OFL_RESOURCE = 1, OFL_STOCKPILE = 2, OFL_HARVEST = 4, OFL_MOVE = 8, OFL_ATTACK = 16 };

typedef struct OUtt OUtt;
typedef struct OGame OGame;
typedef struct OAi OAi;

/* an action as 5 ints: type, parameter, x, y, unitType(-1 = none) */
typedef struct { int type, param, x, y, utype; } OActionV;

OUtt *o_utt_create(int version, int conflict_policy);
OUtt *o_utt_empty(int conflict_policy);
void o_utt_set_type(OUtt *, int id, const int16_t fields[OF_NFIELDS], int flags, int nprod, const uint8_t *prod);
int o_utt_field(const OUtt *, int id, int field);
int o_utt_flags(const OUtt *, int id);
int o_utt_produces(const OUtt *, int id, uint8_t *out);
int o_utt_max_attack_range(const OUtt *);
void o_utt_free(OUtt *);

OGame *o_game_create(const OUtt *, int w, int h, const uint8_t *terrain, int res0, int res1);
void o_game_add_unit(OGame *, int type, int64_t id, int player, int x, int y, int res, int hp);
OGame *o_game_clone(const OGame *);
void o_game_free(OGame *);
void o_game_seed(OGame *, int64_t seed);
int o_game_time(const OGame *);
int o_game_n_units(const OGame *);
int o_game_resources(const OGame *, int player);
int o_game_winner(const OGame *);
int o_game_gameover(const OGame *);
int o_game_errors(const OGame *);
/* raw 48-bit LCG state: 0 policy, 1 conflict, 2 damage */
int64_t o_game_rng_state(const OGame *, int which);
void o_game_set_rng_state(OGame *, int which, int64_t state);
/* out[i*8 .. ] = type, player, x, y, res, hp, id_lo, id_hi  (list order) */
int o_game_units(const OGame *, int32_t *out);
/* per unit in list order: has(0/1), type, param, x, y, utype, issue_time, order (rank in insertion order) */
int o_game_assignments(const OGame *, int32_t *out);
int o_game_cycle(OGame *);
int o_game_is_complete(const OGame *);
int o_game_next_change_time(const OGame *);
/* unit_idx = positions in the current unit list */
int o_game_issue(OGame *, int n, const int32_t *unit_idx, const OActionV *acts, int safe);
int o_game_issue_out(OGame *g, int n, const int32_t *unit_idx, OActionV *acts_inout, int safe);
int o_unit_actions(const OGame *, int unit_idx, int none_duration, OActionV *out, int max_out);
int o_game_free_cell(const OGame *, int x, int y);

/* java.util.Random */
typedef struct { uint64_t s; } OJRandom;
void o_jr_seed(OJRandom *, int64_t seed);
int32_t o_jr_next(OJRandom *, int bits);
int32_t o_jr_next_int(OJRandom *);
int32_t o_jr_next_int_bound(OJRandom *, int32_t bound);
double o_jr_next_double(OJRandom *);

/* policies: return number of (unit_idx, action) pairs written */
int o_ai_random_biased(OGame *, int player, int32_t *unit_idx, OActionV *acts);
OAi *o_ai_create(int kind, int pathfinder);
OAi *o_ai_clone(const OAi *);
void o_ai_free(OAi *);
int o_ai_get_action(OAi *, OGame *, int player, int32_t *unit_idx, OActionV *acts);
/* vector action decode (PlayerAction.fromVectorAction + JNIAI fill NONE(1)) */
int o_from_vector_action(OGame *, int player, int n, const int32_t *vec /*[n][8]*/, int fill_none_duration,
                         int32_t *unit_idx, OActionV *acts);

/* returns direction 0..3 or -1 (null) */
int o_pathfind(const OGame *, int kind, int unit_idx, int targetpos, int range, int n_ru, const int32_t *ru_pos);

/* a FloodFillPathFinding instance (stateful: its cache of distance maps persists across calls) */
void *o_ff_create(void);
void o_ff_free(void *);
int o_ff_find(void *ff, const OGame *, int unit_idx, int targetpos, int range, int n_ru, const int32_t *ru_pos);

/* observations / masks */
void o_observe(const OGame *, int player, int32_t *out /*[6][h][w]*/);
void o_observe_po(const OGame *po_view, int player, int32_t *out /*[8][h][w]*/);
void o_masks(const OGame *, int player, int32_t *out /*[h][w][1+6+16+nTypes+R*R]*/);
OGame *o_po_view(const OGame *, int observer);

/* evaluation: 0 = SimpleSqrtEvaluationFunction3, 1 = SimpleEvaluationFunction */
float o_evaluate(const OGame *, int fn, int maxplayer, int minplayer);

/* loops */
/* Game.start: policies on the same state, issueSafe x2, cycle.  Runs until gameover, time>=max_cycles or
 * n_cycles iterations.  ai0/ai1 may be NULL with kind RANDOM_BIASED/PASSIVE.  Returns 1 if gameover. */
int o_run_game(OGame *, int kind0, OAi *ai0, int kind1, OAi *ai1, int n_cycles, int max_cycles, int64_t *stats);
int o_run_game_po(OGame *, int kind0, OAi *ai0, int kind1, OAi *ai1, int n_cycles, int max_cycles); /* Game(partiallyObservable) */
int o_run_game_observing(OGame *g, int kind0, OAi *ai0, int kind1, OAi *ai1, int n_cycles, int max_cycles, int32_t *scratch);
/* NaiveMCTS.simulate: RandomBiased both sides, issue() not issueSafe() */
int o_simulate(OGame *, int time_limit);

/* PlayerActionGenerator / GameState.getPlayerActions (parity unpinned) */
typedef struct OPag OPag;
OPag *o_pag_create(OGame *, int player, int none_duration); /* the game must outlive the generator; NULL: no unit can act */
void o_pag_free(OPag *);
int64_t o_pag_size(const OPag *);
int64_t o_pag_generated(const OPag *);
int o_pag_n_choices(const OPag *);
int o_pag_next(OPag *, int32_t *unit_idx, OActionV *acts);
void o_pag_randomize_order(OPag *, OJRandom *);
int o_pag_random(OPag *, OJRandom *, int32_t *unit_idx, OActionV *acts);
int64_t o_player_actions(OGame *, int player, int32_t *out, int64_t max_ints);

/* NaiveMCTS with seeded generators (parity unpinned) */
typedef struct OMcts OMcts;
OMcts *o_mcts_create(const OGame *, int player, int lookahead, int max_depth, float e_l, float e_g, float e_0, int strategy, int fensa, int eval_fn, int64_t seed);
void o_mcts_free(OMcts *);
void o_mcts_iterate(OMcts *, int n);
int o_mcts_root(const OMcts *, int *root_visits, double *root_accum, int *out_visits, double *out_accum, int max_children);
int o_mcts_n_nodes(const OMcts *);
int o_mcts_best_action(const OMcts *, int32_t *unit_idx, OActionV *acts);

/* UCT with seeded generators (parity unpinned) */
typedef struct OUct OUct;
OUct *o_uct_create(const OGame *, int player, int lookahead, int max_depth, int eval_fn, int64_t seed);
void o_uct_free(OUct *);
void o_uct_iterate(OUct *, int n);
int o_uct_root(const OUct *, int *root_visits, float *root_accum, int *out_visits, float *out_accum, int max_children);
int o_uct_n_nodes(const OUct *);
int o_uct_best_action(const OUct *, int32_t *unit_idx, OActionV *acts);

#ifdef __cplusplus
}
#endif
#endif
