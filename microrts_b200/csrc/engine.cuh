// engine.cuh -- warp-per-game microRTS rules engine for sm_100a (device code).
//
// One warp owns one game at a time.  The game's unit table (struct-of-arrays, slot order == the reference's unit list
// order) and three byte maps over a wall-padded grid live in shared memory:
//    grid [cell] : 0 empty, 0xFF wall/out of bounds, else unit slot + 1      -> getUnitAt / terrain in O(1)
//    resv [cell] : slot + 1 of the unit whose in-flight MOVE/PRODUCE targets the cell  -> ResourceUsage.positionsUsed
//    claim[cell] : bit p set while player p's policy has tentatively chosen the cell this cycle (PlayerAction.r)
// Lanes work on different units (enumeration, sampling, legality) and arbitrate order-dependent effects with
// ballot / shuffle / redux; effects that the reference applies sequentially are replayed in the same order.
//
// Kernels (step_kernel_body at the end of the file): persistent grids; a warp takes its first game statically and the next ones
// from a global counter.  The hot kernels exist in three forms: with the layout as kernel parameters, as fixed-size template
// copies (MRTS_FIXED_VARIANTS) and, for the generic kernel, as whole translation units compiled for one layout
// (MRTS_TU_FIXED, fixed_generic.inc) -- in the last two the shared-memory offsets are compile-time constants.
//
// Each function cites the reference code it restates (paths under the reference checkout, src/...).
// The file is also compiled for the host by tests/emu (MRTS_EMU) where warp primitives are emulated with
// coroutines; that build is test tooling only and is never linked into libmicrorts_cuda.so.
#pragma once
#include "layout.h"

#ifdef MRTS_EMU
#define DEV static inline
#define DEVN static
#define MDEV inline
#else
#define DEV __device__ __forceinline__
#define MDEV __device__ __forceinline__
#ifdef MRTS_INLINE_ALL
#define DEVN static __device__ __forceinline__
#else
#define DEVN static __device__ __noinline__ // static: the engine is compiled into more than one translation unit (fixed_<W>x<H>.cu)
#endif
#endif

#define FULLM 0xffffffffu
#define MASK48 ((1ULL << 48) - 1)
#define A0_NOUT (0xFFu << 8) // "unitType == null"

enum { ACT_NONE = 0, ACT_MOVE = 1, ACT_HARVEST = 2, ACT_RETURN = 3, ACT_PRODUCE = 4, ACT_ATTACK = 5 };
enum { POL_EXTERNAL = 0, POL_PASSIVE = 1, POL_RANDOM_BIASED = 2, POL_WORKER_RUSH = 3, POL_LIGHT_RUSH = 4, POL_HEAVY_RUSH = 5, POL_RANGED_RUSH = 6,
       POL_WORKER_DEFENSE = 7, POL_LIGHT_DEFENSE = 8, POL_HEAVY_DEFENSE = 9, POL_RANGED_DEFENSE = 10,
       POL_PO_WORKER_RUSH = 11, POL_PO_LIGHT_RUSH = 12, POL_PO_HEAVY_RUSH = 13, POL_PO_RANGED_RUSH = 14, POL_WORKER_RUSH_PP = 15,
       POL_CRUSH_V1 = 16, POL_CRUSH_V2 = 17 /* ai.abstraction.cRush.CRush_V1, CRush_V2 */, POL_EMR_DETERMINISTICO = 18 };
#define POL_IS_SCRIPTED(p) ((p) >= POL_WORKER_RUSH && (p) <= POL_EMR_DETERMINISTICO)
#define POL_IS_PO_RUSH(p) ((p) >= POL_PO_WORKER_RUSH && (p) <= POL_PO_RANGED_RUSH)
#define POL_IS_DEFENSE(p) (((p) >= POL_WORKER_DEFENSE && (p) <= POL_RANGED_DEFENSE) || (p) == POL_WORKER_RUSH_PP) // WorkerRushPlusPlus.java = WorkerDefense.java whose melee units always attack
enum { FMT_VECTOR = 0, FMT_RAW = 1 };
enum { MODE_GAME = 0, MODE_CYCLE_ONLY = 1, MODE_ISSUE_ONLY = 2, MODE_OBSERVE = 3, MODE_MASKS = 4, MODE_ROLLOUT = 5, MODE_PATHFIND = 6,
       MODE_UNIT_ACTIONS = 7, MODE_CYCLE_DECISION = 8 };
enum { ST_OVER = 1, ST_COUNTED = 2 };
enum { STAT_WINS0 = 0, STAT_WINS1, STAT_DRAWS, STAT_FINISHED, STAT_CYCLES, STAT_DECISIONS, STAT_UNIT_CYCLES, STAT_ERRORS,
       STAT_IO_READ, STAT_IO_WRITE, N_WARP_STATS }; // the last two: bytes of game state / outputs the step kernels read from and wrote to global memory (mrts_batch_io_bytes)
#define MRTS_STATS_IO_SLOT 18 // where they live in the batch's counter array: behind the 8 counters, 2 work counters and 8 words of all-reduce scratch

struct StepParams {
    int32_t *hdr;              // [n_games][MRTS_HDR_WORDS]
    uint32_t *units;           // [n_games][7][cap]
    const uint32_t *maps;      // n_maps blobs (layout.h: mrts_map_blob_words)
    const uint32_t *cst;       // MRTS_CONST_WORDS: unit type table + LCG jump table
    unsigned long long *stats; // [8]
    int32_t *results_out;      // MODE_GAME: [n_games][4] {time, winner, gameover, error bits} of the state each game is left in (or NULL)
    unsigned long long *work_counter; // [2]: items handed out beyond each warp's first one; warps that have left the launch.
                                      // Both are zero between launches: the last warp to leave resets them (no memset per launch,
                                      // which would queue behind other streams' copies)
    long long n_games;
    int n_maps, map_words;
    int W, H, cap;
    SmemLayout L;              // host-computed (mrts_smem_layout): offsets reach the kernels through the constant bank
    int mode;
    int n_cycles, max_cycles;
    const int32_t *t_target;   // MODE_CYCLE_ONLY: absolute target time per game (or NULL)
    int policy[2], pathfinder[2];
    int conflict;              // UnitTypeTable.moveConflictResolutionStrategy
    int safe;                  // issueSafe (1) or issue (0) for external actions
    int issue_player;          // MODE_ISSUE_ONLY
    int scripted;              // 0: no scripted policies; 1: A* scratch in shared memory; 2: A* scratch in global memory
    int uw;                    // unit words per slot in HBM and shared memory: 7, or 9 (X0/X1) for scripted batches
    unsigned char *astar_scratch; // scripted == 2: [grid * warps per CTA][astar_stride] bytes
    long long astar_stride;
    unsigned char *ff_cache;      // MRTS_PF_FLOODFILL: [n_games][2][ff_stride] bytes (scripted.cuh: pf_floodfill), or null
    long long ff_stride;
    int auto_reset;            // MODE_GAME: restart finished games from their map at the start of the step
    const int32_t *ext_actions[2]; // [n_games][max_k][8]
    const int32_t *ext_counts[2];  // [n_games]
    int ext_maxk[2], ext_format[2], ext_fill[2];
    long long ext_stride[2];   // int32 elements between consecutive games' rows (max_k * 8, or twice that for an interleaved array)
    int vec_reset;             // MODE_GAME, JNIGridnetVecClient.gameStep's auto-reset inside the launch: 0 off; 1 restart when the game is over; 2 when no
                               // Resource unit holds resources any more; 3 never on the state -- always also after vec_max_steps steps
    int vec_max_steps;
    int out_stride;            // fused outputs (obs_out, mask_out): game g is written at game slot g * out_stride of the buffers
    int max_range;             // UnitTypeTable.getMaxAttackRange()
    int n_types;               // utt.getUnitTypes().size()
    // MODE_PATHFIND: one query per game -- {cell of the unit (x + y*W), target position (x + y*W), range} -> direction or -1
    const int32_t *pf_query;   // [n_games][3]
    int32_t *pf_out;           // [n_games]
    int pf_kind;
    // MODE_UNIT_ACTIONS: the ordered legal action lists of out_player's idle units + the state's resource usage (mrts_batch_unit_actions)
    int32_t *ua_hdr, *ua_pos, *ua_choice, *ua_list; // [n][8], [n][cap], [n][K][4], [n][K][ua_max_actions] with K = ua_max_choices
    int ua_max_choices, ua_max_actions, ua_none_duration;
    // MODE_OBSERVE / MODE_MASKS
    void *out;
    int out_dtype;             // 0 = u8, 1 = i32, 2 = bit-packed (masks only)
    int out_player;
    int partial_obs;
    int po_policies;           // MRTS_FLAG_PO_POLICIES: device policies decide on their player's partially observable view
    int sequential_issue;      // MODE_GAME: player 1 decides on the state that already holds player 0's actions
                               // (JNIGridnetClientSelfPlay.gameStep) instead of both deciding on the pre-issue state (Game.start)
    int32_t *info_out;         // MODE_GAME: [n_games][2][MRTS_INFO_WORDS] per-player step facts for the reward functions (or NULL)
    uint32_t tm_worker, tm_building, tm_combat, tm_base, tm_mobile, tm_resource; // unit type id bit masks, resolved by name on the host
    void *obs_out[2];          // MODE_GAME: when set, the post-step observation of player 0 / 1 is written here ([n][6][H][W])
    int obs_dtype;
    void *mask_out[2];         // MODE_GAME: when set, the post-step bit-packed action masks of player 0 / 1 ([n][H][W][(K+7)/8] bytes)
    int zero_bytes;            // KERNEL_FAST_OBS: size of the CTA's block of zeros behind the warps' regions (source of the bulk stores), or 0
    int terr_bytes, tmpl_bytes; // ... followed by the map's terrain plane (source of plane 5's bulk store) and its wall-padded grid template (single-map batches)
    // MODE_ROLLOUT: item r = game r / rollouts_per_game
    int rollouts_per_game, depth, eval_fn, maxplayer, observer;
    const long long *ro_seeds; // [n_games * rollouts_per_game] or NULL (seed = r)
    float *ro_eval;            // [n_games * rollouts_per_game]
    int32_t *ro_time;          // [n_games * rollouts_per_game] simulated cycles
};

// The CTA's dynamic shared memory.  A game's region and the constant block are addressed through 32-bit shared-window
// addresses that are computed once per thread and kept in registers (smem_window / smem_ptr): the compiler infers the
// address space from the conversion, so every access is an LDS/STS with 32-bit addressing even inside functions that are
// not inlined, and it does not re-derive the window base from the special registers at every use.
#ifdef MRTS_EMU
static thread_local unsigned char *mrts_smem = nullptr;
static inline uint32_t smem_window(int byte_offset) { return (uint32_t)byte_offset; }
static inline unsigned char *smem_ptr(uint32_t a) { return mrts_smem + a; }
#else
extern __shared__ __align__(16) unsigned char mrts_smem[];
__device__ __forceinline__ uint32_t smem_window(int byte_offset) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(mrts_smem) + (uint32_t)byte_offset;
    asm volatile("" : "+r"(a)); // opaque: stays in a register instead of being rematerialised
    return a;
}
__device__ __forceinline__ unsigned char *smem_ptr(uint32_t a) { return (unsigned char *)__cvta_shared_to_generic(a); }
#endif

// MRTS_TU_FIXED: this whole translation unit is compiled for ONE layout (map MRTS_TU_W x MRTS_TU_H, MRTS_TU_CAP unit slots,
// scripted-policy words and pathfinding scratch in shared memory; fixed_generic.inc).  The layout fields of Game are then static
// constants, so they fold into immediates even inside the out-of-line functions of the generic kernel, which otherwise
// reload them from the Game object in local memory at every access.
#ifdef MRTS_TU_FIXED
#define MRTS_TU_LAYOUT mrts_smem_layout(MRTS_TU_W, MRTS_TU_H, MRTS_TU_CAP, 1, 0, 1)
#endif
struct Game {
    int lane;
#ifdef MRTS_TU_FIXED
    static constexpr int W = MRTS_TU_W, H = MRTS_TU_H, P = MRTS_TU_W + 2, cap = MRTS_TU_CAP, pcw = MRTS_TU_LAYOUT.pcw, uw = MRTS_TU_LAYOUT.uws - 1;
    static constexpr int o_pa0 = MRTS_TU_LAYOUT.pa0, o_pa1 = MRTS_TU_LAYOUT.pa1, o_pslot = MRTS_TU_LAYOUT.pslot, o_grid = MRTS_TU_LAYOUT.grid,
                         o_kind = MRTS_TU_LAYOUT.kind, o_resv = MRTS_TU_LAYOUT.resv, o_claim = MRTS_TU_LAYOUT.claim, o_list = MRTS_TU_LAYOUT.list,
                         o_povis = MRTS_TU_LAYOUT.povis, o_pohid = MRTS_TU_LAYOUT.pohid, o_rdy = MRTS_TU_LAYOUT.rdy, o_units = MRTS_TU_LAYOUT.uoff[0];
    int conflict;
    static constexpr bool slim = false;
    MDEV static constexpr int uoffset(int k) { return o_units + k * cap * 4; }
#else
    bool slim;                            // no kind / claim maps (layout.h: the fused step + observation kernel)
    int W, H, P, cap, pcw, conflict, uw; // uw: unit words mirrored in HBM (7, or 9 with scripted policies)
    int o_pa0, o_pa1, o_pslot, o_grid, o_kind, o_resv, o_claim, o_list;
    int o_povis, o_pohid;                 // MRTS_FLAG_PO_POLICIES batches only (layout.h)
    int o_rdy, uoff[MRTS_UNIT_WORDS + 1]; // byte offsets of the unit word arrays inside the region (host-computed constants)
    MDEV int uoffset(int k) const { return uoff[k]; }
#endif
    uint32_t sb, cb;                      // shared-window addresses of this game's region and of the CTA's constant block
    bool po_view;                         // the unit table currently shows one player's partially observable view (po_hide)
    int pview;                            // window into the pending list (policy_scripted stages desires behind the final part)
    const uint32_t *grid_tmpl;            // global: wall-padded empty grid of this game's map
    uint16_t *as_mark, *as_next, *as_head, *as_gen; // A*/BFS scratch of this warp (scripted batches only, layout.h)
    uint32_t as_sm;                       // its shared-window address when it lives in shared memory, else 0
    const uint4 *tmpl_sm;                 // the CTA's copy of the wall-padded empty grid in shared memory (fused step + observation kernel), or null
    unsigned char *ff_cache;              // FloodFillPathFinding: this game's two per-player distance-map caches in HBM (or null)
    long long ff_stride;                  // bytes per player

    MDEV unsigned char *base() const { return smem_ptr(sb); }
    MDEV int32_t *hdr() const { return (int32_t *)base(); }
    MDEV uint32_t *uword(int k) const { return (uint32_t *)(base() + uoffset(k)); }
    MDEV uint32_t *w0() const { return uword(UW_W0); }
    MDEV uint32_t *w1() const { return uword(UW_W1); }
    MDEV uint32_t *a0() const { return uword(UW_A0); }
    MDEV int32_t *a1() const { return (int32_t *)uword(UW_A1); }
    MDEV int32_t *tis() const { return (int32_t *)uword(UW_TIS); }
    MDEV uint32_t *seq() const { return uword(UW_SEQ); }
    MDEV uint32_t *uid() const { return uword(UW_ID); }
    MDEV uint32_t *x0() const { return uword(UW_X0); }
    MDEV uint32_t *x1() const { return uword(UW_X1); }
    MDEV int32_t *rdy() const { return (int32_t *)(base() + o_rdy); } // shared memory only: completion time, MRTS_NEVER when idle
    MDEV uint32_t *pa0() const { return (uint32_t *)(base() + o_pa0) + pview; }
    MDEV int32_t *pa1() const { return (int32_t *)(base() + o_pa1) + pview; }
    MDEV uint8_t *pslot() const { return base() + o_pslot + pview; }
    MDEV uint8_t *grid() const { return base() + o_grid; }
    MDEV uint8_t *kind() const { return base() + o_kind; }
    MDEV uint8_t *resv() const { return base() + o_resv; }
    MDEV uint8_t *claim() const { return base() + o_claim; }
    MDEV uint8_t *list() const { return base() + o_list; }
    MDEV uint8_t *vis() const { return base() + o_povis; }
    MDEV uint32_t *hid() const { return (uint32_t *)(base() + o_pohid); }
    MDEV const uint32_t *utt() const { return (const uint32_t *)smem_ptr(cb); }
    MDEV const uint64_t *jump() const { return (const uint64_t *)smem_ptr(cb + MRTS_MAX_TYPES * MRTS_UTT_WORDS * 4); }
};

DEV void g_bind(Game &g, int region, const SmemLayout &L, int W, int H, int cap, int lane, int conflict, int scripted,
                unsigned char *astar_global, bool slim = false) {
    g.lane = lane; g.conflict = conflict;
#ifndef MRTS_TU_FIXED
    g.slim = slim;
    g.W = W; g.H = H; g.P = L.P; g.cap = cap; g.pcw = L.pcw;
    for (int k = 0; k <= MRTS_UNIT_WORDS; k++) g.uoff[k] = L.uoff[k];
    g.o_rdy = L.rdy;
    g.uw = L.uws - 1; // host-computed: 7, or 9 for scripted batches
    g.o_pa0 = L.pa0; g.o_pa1 = L.pa1; g.o_pslot = L.pslot; g.o_grid = L.grid; g.o_kind = L.kind; g.o_resv = L.resv;
    g.o_claim = L.claim; g.o_list = L.list; g.o_povis = L.povis; g.o_pohid = L.pohid;
#endif
    g.sb = smem_window(region); g.cb = smem_window(0); g.pview = 0; g.po_view = false;
    { int pc = (W + 2) * (H + 2); g.as_mark = (uint16_t *)(astar_global ? astar_global : mrts_smem + region + L.astar);
      g.as_next = g.as_mark + pc; g.as_head = g.as_next + pc; g.as_gen = g.as_head + MRTS_ASTAR_HEADS(W, H);
      g.as_sm = (scripted == 1) ? smem_window(region + L.astar) : 0u; }
    g.grid_tmpl = nullptr; g.ff_cache = nullptr; g.ff_stride = 0; g.tmpl_sm = nullptr;
}

// ---- small accessors -------------------------------------------------------------------------------------------------
DEV int u_type(uint32_t w) { return w & 0xff; }
DEV int u_pl(uint32_t w) { return (w >> 8) & 0xff; } // 0 neutral, 1 = player 0, 2 = player 1
#define HIDDEN_TYPE 7 // po_hide: a unit hidden from the deciding player's view shows as a neutral unit of this unused type
DEV bool w_hidden(uint32_t w) { return (w & 0xffffu) == (uint32_t)HIDDEN_TYPE; }
DEV int u_x(uint32_t w) { return (w >> 16) & 0xff; }
DEV int u_y(uint32_t w) { return w >> 24; }
DEV int cell_of(const Game &g, uint32_t w) { return (u_y(w) + 1) * g.P + u_x(w) + 1; }
DEV int doff(const Game &g, int d) { return (d & 1) ? 2 - d : (d - 1) * g.P; } // up -P, right +1, down +P, left -1 (UnitAction.java:94,100)
DEV int ddx(int d) { return d == 1 ? 1 : (d == 3 ? -1 : 0); }
DEV int ddy(int d) { return d == 0 ? -1 : (d == 2 ? 1 : 0); }
DEV int ut_cost(const Game &g, int t) { return g.utt()[t * 8] & 0xff; }
DEV int ut_hp(const Game &g, int t) { return (g.utt()[t * 8] >> 8) & 0xff; }
DEV int ut_mind(const Game &g, int t) { return (g.utt()[t * 8] >> 16) & 0xff; }
DEV int ut_maxd(const Game &g, int t) { return g.utt()[t * 8] >> 24; }
DEV int ut_range(const Game &g, int t) { return g.utt()[t * 8 + 1] & 0xff; }
DEV int ut_sight(const Game &g, int t) { return (g.utt()[t * 8 + 1] >> 8) & 0xff; }
DEV int ut_hamt(const Game &g, int t) { return (g.utt()[t * 8 + 1] >> 16) & 0xff; }
DEV int ut_flags(const Game &g, int t) { return g.utt()[t * 8 + 1] >> 24; }
DEV int ut_nprod(const Game &g, int t) { return (g.utt()[t * 8 + 4] >> 16) & 0xff; }
DEV int ut_prod(const Game &g, int t, int k) { return (g.utt()[t * 8 + 5 + (k >> 2)] >> ((k & 3) * 8)) & 0xff; }
DEV int a_type(uint32_t A0) { return A0 & 0xF; }
DEV int a_utype(uint32_t A0) { return (A0 >> 8) & 0xff; }
DEV bool a_uses_cell(int at) { return at == ACT_MOVE || at == ACT_PRODUCE; }
DEV int u_hp(uint32_t w1) { return (int)(int16_t)(w1 & 0xffff); }
DEV int u_res(uint32_t w1) { return (int)(int16_t)(w1 >> 16); }
DEV uint32_t mk_w1(int hp, int res) { return ((uint32_t)hp & 0xffffu) | ((uint32_t)res << 16); }
DEV int nth4(int m, int n) { // index of the n-th set bit of a 4-bit mask: byte m of the table holds the positions, 2 bits each
    unsigned long long T = (m & 8) ? 0xe439380e340d0c03ULL : 0x2409080204010000ULL;
    return (int)(T >> ((m & 7) * 8 + 2 * n)) & 3;
}
DEV int nth8(int m, int n) {
    #pragma unroll 1
    for (int d = 0; d < 8; d++) {
        if ((m >> d) & 1) { if (n == 0) return d; n--; }
    }
    return 0;
}

// UnitAction.ETA, UnitAction.java:307-329 (RETURN uses moveTime; PRODUCE the produced type's produceTime): one lookup in
// the u16 table eta[type][action type] behind the unit type table
DEV int eta_of(const Game &g, int t, uint32_t A0, int A1) {
    int at = a_type(A0);
    if (at == ACT_NONE) return A1;
    if (at > ACT_ATTACK) return 0;
    if (at == ACT_PRODUCE) { t = a_utype(A0); if (t >= MRTS_MAX_TYPES) return 0; }
    return ((const uint16_t *)(g.utt() + MRTS_ETA_OFFSET))[t * 8 + at];
}
// target cell of a MOVE/PRODUCE (UnitAction.resourceUsage, UnitAction.java:254-291; an out-of-range direction leaves
// `pos` at the unit's own cell)
DEV int target_cell(const Game &g, int c, int A1) { return ((unsigned)A1 < 4u) ? c + doff(g, A1) : c; }

// The reference computes the used position with linear arithmetic, pos = x + y*W -/+ {W,1} (UnitAction.java:255-270), so
// an (illegal) move off the left/right edge aliases the neighbouring row's end cell.  Returns the padded-grid index of
// that linear position, or -1 when it falls outside the map.  Only external (possibly illegal) actions need this.
DEV int linear_target_cell(const Game &g, uint32_t w, int A1) {
    int lin = u_x(w) + u_y(w) * g.W;
    if ((unsigned)A1 < 4u) lin += (A1 == 0 ? -g.W : (A1 == 1 ? 1 : (A1 == 2 ? g.W : -1)));
    if (lin < 0 || lin >= g.W * g.H) return -1;
    return (lin / g.W + 1) * g.P + lin % g.W + 1;
}

// ---- java.util.Random (Java SE spec: 48-bit LCG) -----------------------------------------------------------------------
DEV uint64_t lcg_next(uint64_t s) { return (s * 0x5DEECE66DULL + 0xBULL) & MASK48; }
DEV uint64_t lcg_jump(const Game &g, uint64_t s, int draws) { // advance by `draws` nextDouble() calls (2 steps each)
    #pragma unroll 1
    while (draws > 64) { s = (s * g.jump()[128] + g.jump()[129]) & MASK48; draws -= 64; }
    return (s * g.jump()[2 * draws] + g.jump()[2 * draws + 1]) & MASK48;
}
DEV double lcg_next_double(uint64_t &s) { // ((long)next(26) << 27) + next(27)) * 2^-53
    s = lcg_next(s); long long a = (long long)(s >> 22);
    s = lcg_next(s); long long b = (long long)(s >> 21);
    return (double)((a << 27) + b) * 0x1.0p-53;
}
DEV int lcg_next_int_bound(uint64_t &s, int bound) { // Random.nextInt(bound)
    s = lcg_next(s);
    int r = (int)(s >> 17);
    int m = bound - 1;
    if ((bound & m) == 0) return (int)(((long long)bound * (long long)r) >> 31);
    #pragma unroll 1
    for (int u = r;;) {
        r = u % bound;
        if ((int)((unsigned)u - (unsigned)r + (unsigned)m) >= 0) break;
        s = lcg_next(s); u = (int)(s >> 17);
    }
    return r;
}
DEV uint64_t hdr_rng(const Game &g, int lo) { return (uint64_t)(uint32_t)g.hdr()[lo] | ((uint64_t)(uint32_t)g.hdr()[lo + 1] << 32); }
DEV void hdr_set_rng(Game &g, int lo, uint64_t s) { g.hdr()[lo] = (int32_t)(uint32_t)s; g.hdr()[lo + 1] = (int32_t)(uint32_t)(s >> 32); }

// ---- load / store ------------------------------------------------------------------------------------------------------
DEV int kind_of(const Game &g, uint32_t w) { // the cell-kind byte of a unit (layout.h)
    int fl = ut_flags(g, u_type(w));
    return CK_UNIT | u_pl(w) | ((fl & UF_RESOURCE) ? 4 : 0) | ((fl & UF_STOCKPILE) ? 8 : 0);
}
// Rebuild the cell maps from the unit table: grid (slot+1), kind, resv (in-flight MOVE/PRODUCE targets); claim cleared.
// with_rdy: also recompute the completion times (after a load; compaction carries them along instead).
// 16-byte asynchronous copies global -> shared (LDGSTS): a lane moves four unit slots of one word array per instruction and all of
// a game's copies are in flight together, so loading a game costs one memory round trip instead of one per word array.
#ifdef MRTS_EMU
DEV void cp_async16(void *smem_dst, const void *gsrc) { memcpy(smem_dst, gsrc, 16); }
DEV void cp_async_wait_all() {}
#else
DEV void cp_async16(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
DEV void cp_async_wait_all() { asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory"); }
#endif
// the wall-padded empty grid of the game's map into grid[] and kind[] (asynchronous 16-byte copies: all in flight together, and
// together with the unit words when a game is loaded), zeros into resv[] and claim[]
DEV void g_template(Game &g) {
    uint4 z; z.x = z.y = z.z = z.w = 0;
    int nq = g.pcw >> 2;
    if (g.tmpl_sm) { // the CTA keeps the template in shared memory: no L2 traffic per game
        #pragma unroll 2
        for (int i = g.lane; i < nq; i += 32) {
            uint4 t = g.tmpl_sm[i];
            ((uint4 *)g.grid())[i] = t;
            ((uint4 *)g.resv())[i] = z;
            if (!g.slim) { ((uint4 *)g.kind())[i] = t; ((uint4 *)g.claim())[i] = z; }
        }
        return;
    }
    #pragma unroll 2
    for (int i = g.lane; i < nq; i += 32) {
        cp_async16((uint4 *)g.grid() + i, (const uint4 *)g.grid_tmpl + i);
        ((uint4 *)g.resv())[i] = z;
        if (!g.slim) {
            cp_async16((uint4 *)g.kind() + i, (const uint4 *)g.grid_tmpl + i);
            ((uint4 *)g.claim())[i] = z;
        }
    }
}
DEV void g_rebuild(Game &g, bool with_rdy, bool template_pending = false) {
    __syncwarp();
    if (!template_pending) g_template(g);
    cp_async_wait_all();
    __syncwarp();
    int n = g.hdr()[H_NUNITS];
    #pragma unroll 1
    for (int i = g.lane; i < n; i += 32) {
        uint32_t w = g.w0()[i];
        int c = cell_of(g, w);
        g.grid()[c] = (uint8_t)(i + 1);
        if (!g.slim) g.kind()[c] = (uint8_t)kind_of(g, w);
        uint32_t A0 = g.a0()[i];
        int A1 = g.a1()[i];
        if (a_uses_cell(a_type(A0))) g.resv()[target_cell(g, c, A1)] = (uint8_t)(i + 1);
        if (with_rdy) g.rdy()[i] = a_type(A0) == (int)AT_IDLE ? MRTS_NEVER : g.tis()[i] + eta_of(g, u_type(w), A0, A1);
    }
    __syncwarp();
}
// unit words [0, uw) of slots [0, n) from `src` ([uw][cap] words, 16-byte aligned: cap is a multiple of 4) into the game's table;
// slots up to the next multiple of 4 come along (never read: every pass over the table stops at nUnits)
DEV void load_unit_words(Game &g, const uint32_t *src, int n) {
    uint32_t *su = g.w0();
    #pragma unroll 1
    for (int i = g.lane * 4; i < n; i += 128)
        #pragma unroll
        for (int k = 0; k < MRTS_UNIT_WORDS; k++) if (k < g.uw) cp_async16(su + k * g.cap + i, src + k * g.cap + i);
}
// Load the game from HBM; when `restart_if_over` and the game ended in an earlier step (game over, or time >= max_cycles),
// start it again from the map's initial state instead (the blob's init header/units): the three RNG streams keep running,
// as the reference's static Random objects do across games, and H_SPARE counts episodes.
// The first 128 unit slots are requested together with the header (before the unit count is known): one round trip for almost
// every game; only games with more live units, or restarted ones, pay a second.
DEV void g_load(Game &g, const int32_t *ghdr, const uint32_t *gun, bool restart_if_over, int max_cycles) {
    __syncwarp();
    int32_t v = 0;
    if (g.lane < MRTS_HDR_WORDS) v = ghdr[g.lane];
    const int spec = g.cap < 128 ? g.cap : 128; // speculative part
    load_unit_words(g, gun, spec);
    g_template(g); // the cell maps' template travels with them
    bool restart = false;
    if (restart_if_over) {
        int st = __shfl_sync(FULLM, v, H_STATUS), tm = __shfl_sync(FULLM, v, H_TIME);
        restart = (st & ST_OVER) || tm >= max_cycles;
    }
    if (restart) {
        if (g.lane < MRTS_HDR_WORDS) {
            int32_t iv = ((const int32_t *)(g.grid_tmpl + g.pcw))[g.lane];
            if (g.lane == H_SPARE) iv = v + 1;
            if (!(g.lane >= H_RNGP_LO && g.lane <= H_RNGD_HI)) v = iv;
        }
    }
    if (g.lane < MRTS_HDR_WORDS) g.hdr()[g.lane] = v;
    int n = __shfl_sync(FULLM, v, H_NUNITS);
    if (restart) {
        cp_async_wait_all(); // the speculative copies have landed before the same words are requested from the map
        __syncwarp();
        load_unit_words(g, g.grid_tmpl + g.pcw + MRTS_HDR_WORDS, n);
    } else if (n > spec) {
        uint32_t *su = g.w0();
        #pragma unroll 1
        for (int i = spec + g.lane * 4; i < n; i += 128)
            #pragma unroll
            for (int k = 0; k < MRTS_UNIT_WORDS; k++) if (k < g.uw) cp_async16(su + k * g.cap + i, gun + k * g.cap + i);
    }
    g_rebuild(g, true, true); // waits for every copy above
}
DEV void g_store(Game &g, int32_t *ghdr, uint32_t *gun) {
    __syncwarp();
    if (g.lane < MRTS_HDR_WORDS) ghdr[g.lane] = g.hdr()[g.lane];
    int n = g.hdr()[H_NUNITS];
    const uint32_t *su = g.w0();
    #pragma unroll 1
    for (int i = g.lane * 4; i < n; i += 128) // 16 bytes (four slots of one word array) per lane and store
        #pragma unroll
        for (int k = 0; k < MRTS_UNIT_WORDS; k++) if (k < g.uw) *(uint4 *)(gun + k * g.cap + i) = *(const uint4 *)(su + k * g.cap + i);
    __syncwarp();
}

// ---- Unit.getUnitActions (units/Unit.java:382-522) as category masks ----------------------------------------------------
struct Enum {
    uint32_t w;
    int t, pl, c, fl, range;
    int free_m, atk_m, harv_m, ret_m, aff_m;
    int n_atk, nfree, n_aff, nb, total;
};

DEV bool enemy_in_range(const Game &g, uint32_t me, uint32_t ow, int sq) {
    int opl = u_pl(ow);
    if (opl == 0 || opl == u_pl(me)) return false;
    int dx = u_x(ow) - u_x(me), dy = u_y(ow) - u_y(me);
    return dx * dx + dy * dy <= sq;
}

DEV void enumerate(const Game &g, int s, Enum &e) {
    uint32_t w = g.w0()[s];
    e.w = w; e.t = u_type(w); e.pl = u_pl(w); e.c = cell_of(g, w);
    e.fl = ut_flags(g, e.t); e.range = ut_range(g, e.t);
    int myres = u_res(g.w1()[s]);
    // The four neighbours' cell-kind bytes (up, right, down, left) packed into one word and classified for all four
    // directions at once.  Kind byte (layout.h): 0 empty, 0xFF wall, else 0x10 | owner (bits 0-1) | resource << 2 | stockpile << 3.
    uint32_t K;
    if (!g.slim) {
        const uint8_t *kind = g.kind() + e.c;
        K = (uint32_t)kind[-g.P] | ((uint32_t)kind[1] << 8) | ((uint32_t)kind[g.P] << 16) | ((uint32_t)kind[-1] << 24);
    } else { // no kind map: the occupant's kind through grid[] and the unit table
        const uint8_t *gr = g.grid() + e.c;
        int g0 = gr[-g.P], g1 = gr[1], g2 = gr[g.P], g3 = gr[-1];
        uint32_t k0 = (g0 == 0 || g0 == 0xFF) ? (uint32_t)g0 : (uint32_t)kind_of(g, g.w0()[g0 - 1]), k1 = (g1 == 0 || g1 == 0xFF) ? (uint32_t)g1 : (uint32_t)kind_of(g, g.w0()[g1 - 1]);
        uint32_t k2 = (g2 == 0 || g2 == 0xFF) ? (uint32_t)g2 : (uint32_t)kind_of(g, g.w0()[g2 - 1]), k3 = (g3 == 0 || g3 == 0xFF) ? (uint32_t)g3 : (uint32_t)kind_of(g, g.w0()[g3 - 1]);
        K = k0 | (k1 << 8) | (k2 << 16) | (k3 << 24);
    }
    const uint32_t LSB = 0x01010101u;
    uint32_t unit = (K >> 4) & ~(K >> 7) & LSB;                       // occupied by a unit (walls have bit 7)
    uint32_t fre = ~((K >> 4) | (K >> 7)) & LSB;                      // neither unit nor wall
    uint32_t p0 = K & unit, p1 = (K >> 1) & unit;                     // owner is player 0 / player 1
    uint32_t own = e.pl == 1 ? p0 : (e.pl == 2 ? p1 : 0u), enemy = e.pl == 1 ? p1 : (e.pl == 2 ? p0 : 0u);
    // byte d -> bit d: the multiply gathers bits 0, 8, 16, 24 into bits 24..27
    int free_m = (int)((fre * 0x01020408u) >> 24) & 0xF;
    int atk_m = (int)((enemy * 0x01020408u) >> 24) & 0xF;             // an opponent's unit (Unit.java:411-423)
    int harv_m = (int)((((K >> 2) & unit) * 0x01020408u) >> 24) & 0xF; // a resource (Unit.java:440-452)
    int ret_m = (int)((((K >> 3) & own) * 0x01020408u) >> 24) & 0xF;   // an own stockpile (Unit.java:453-465)
    if (!(e.fl & UF_ATTACK) || e.range != 1) atk_m = 0;
    if (!(e.fl & UF_HARVEST)) { harv_m = 0; ret_m = 0; }
    else { if (myres != 0) harv_m = 0; if (!(myres > 0)) ret_m = 0; }
    int n_atk = __popc(atk_m);
    if ((e.fl & UF_ATTACK) && e.range > 1) { // every enemy within range, in unit-list order (Unit.java:424-436)
        int n = g.hdr()[H_NUNITS], sq = e.range * e.range;
        n_atk = 0;
        #pragma unroll 1
        for (int i = 0; i < n; i++) n_atk += enemy_in_range(g, w, g.w0()[i], sq) ? 1 : 0;
    }
    int aff_m = 0;
    if (e.pl != 0) {
        int pres = g.hdr()[H_RES0 + e.pl - 1], np = ut_nprod(g, e.t);
        uint32_t pc = g.utt()[e.t * 8 + 7]; // costs of the first four produced types, one byte each
        aff_m = ((pres >= (int)(pc & 0xff)) ? 1 : 0) | ((pres >= (int)((pc >> 8) & 0xff)) ? 2 : 0) | ((pres >= (int)((pc >> 16) & 0xff)) ? 4 : 0) |
                ((pres >= (int)(pc >> 24)) ? 8 : 0);
        #pragma unroll 1
        for (int k = 4; k < np; k++) if (pres >= ut_cost(g, ut_prod(g, e.t, k))) aff_m |= 1 << k;
        aff_m &= (1 << np) - 1;
    }
    e.free_m = free_m; e.atk_m = atk_m; e.harv_m = harv_m; e.ret_m = ret_m; e.aff_m = aff_m;
    e.n_atk = n_atk; e.nfree = __popc(free_m); e.n_aff = __popc(aff_m);
    e.nb = n_atk + __popc(harv_m) + __popc(ret_m);
    e.total = 5 * e.nb + e.nfree * (e.n_aff + ((e.fl & UF_MOVE) ? 1 : 0)) + 1;
}

// the idx-th action of the list getUnitActions would return; out: packed action, its target cell (-1) and cost
DEV void pick_action(const Game &g, const Enum &e, int idx, int none_duration, uint32_t &A0, int &A1, int &tcell, int &cost) {
    tcell = -1; cost = 0;
    int x = u_x(e.w), y = u_y(e.w);
    if (idx < e.n_atk) {
        int ax, ay;
        if (e.range == 1) { int d = nth4(e.atk_m, idx); ax = x + ddx(d); ay = y + ddy(d); }
        else {
            int n = g.hdr()[H_NUNITS], sq = e.range * e.range; ax = x; ay = y;
            #pragma unroll 1
            for (int i = 0; i < n; i++) {
                uint32_t ow = g.w0()[i];
                if (enemy_in_range(g, e.w, ow, sq)) { if (idx == 0) { ax = u_x(ow); ay = u_y(ow); break; } idx--; }
            }
        }
        A0 = ACT_ATTACK | A0_NOUT | ((uint32_t)ax << 16) | ((uint32_t)ay << 24); A1 = -1; return;
    }
    idx -= e.n_atk;
    int nh = __popc(e.harv_m);
    if (idx < nh) { A0 = ACT_HARVEST | A0_NOUT; A1 = nth4(e.harv_m, idx); return; }
    idx -= nh;
    int nr = __popc(e.ret_m);
    if (idx < nr) { A0 = ACT_RETURN | A0_NOUT; A1 = nth4(e.ret_m, idx); return; }
    idx -= nr;
    int np = e.nfree * e.n_aff;
    if (idx < np) {
        int k = idx / e.nfree, d = nth4(e.free_m, idx - k * e.nfree);
        int ut = ut_prod(g, e.t, nth8(e.aff_m, k));
        A0 = ACT_PRODUCE | ((uint32_t)ut << 8); A1 = d; tcell = e.c + doff(g, d); cost = ut_cost(g, ut); return;
    }
    idx -= np;
    if ((e.fl & UF_MOVE) && idx < e.nfree) { int d = nth4(e.free_m, idx); A0 = ACT_MOVE | A0_NOUT; A1 = d; tcell = e.c + doff(g, d); return; }
    A0 = ACT_NONE | A0_NOUT; A1 = none_duration;
}

// Sampler.weighted (util/Sampler.java:116-137) over weights {5 x nb, 1 x rest}: first i with accum_i >= tmp.
// accum_i are exact integers, so accum_i >= tmp  <=>  accum_i >= ceil(tmp).
DEV int sample_index(double draw, int nb, int total) {
    double tmp = __dmul_rn(draw, (double)total);
    int c = (int)ceil(tmp);
    if (5 * nb >= c) { int j = (c + 4) / 5; return j < 1 ? 0 : j - 1; }
    return nb + (c - 5 * nb) - 1;
}

// Unit.canExecuteAction (units/Unit.java:531-534): is (A0,A1) in getUnitActions(u, ETA)?  NONE is always legal.
DEV bool action_is_legal(const Game &g, int s, uint32_t A0, int A1) {
    int at = a_type(A0);
    bool dir_ok = (unsigned)A1 < 4u;
    if (at == ACT_NONE) return true;
    if (at == ACT_MOVE && !g.slim) { // the common cases without the full enumeration: the target cell's kind byte decides a MOVE
        if (!dir_ok) return false;
        uint32_t w = g.w0()[s];
        int k = g.kind()[cell_of(g, w) + doff(g, A1)];
        return (ut_flags(g, u_type(w)) & UF_MOVE) && !((k >> 4) & 1) && !((k >> 7) & 1); // neither unit nor wall, as enumerate's free_m
    }
    if (at == ACT_ATTACK) { // the attacked cell holds an enemy within range (the enumeration would walk the whole unit list for a ranged unit)
        uint32_t w = g.w0()[s];
        int t = u_type(w), range = ut_range(g, t);
        if (!(ut_flags(g, t) & UF_ATTACK)) return false;
        int ax = (A0 >> 16) & 0xff, ay = A0 >> 24;
        if (ax >= g.W || ay >= g.H) return false;
        int gv = g.grid()[(ay + 1) * g.P + ax + 1];
        if (gv == 0 || gv == 0xFF) return false;
        return enemy_in_range(g, w, g.w0()[gv - 1], range * range);
    }
    Enum e; enumerate(g, s, e);
    switch (at) {
        case ACT_NONE: return true;
        case ACT_MOVE: return dir_ok && (e.fl & UF_MOVE) && ((e.free_m >> A1) & 1);
        case ACT_HARVEST: return dir_ok && ((e.harv_m >> A1) & 1);
        case ACT_RETURN: return dir_ok && ((e.ret_m >> A1) & 1);
        case ACT_PRODUCE: {
            if (!dir_ok || !((e.free_m >> A1) & 1)) return false;
            int ut = a_utype(A0), np = ut_nprod(g, e.t);
            #pragma unroll 1
            for (int k = 0; k < np; k++) if (ut_prod(g, e.t, k) == ut && ((e.aff_m >> k) & 1)) return true;
            return false;
        }
        case ACT_ATTACK: {
            if (!(e.fl & UF_ATTACK)) return false;
            int ax = (A0 >> 16) & 0xff, ay = A0 >> 24;
            if (ax >= g.W || ay >= g.H) return false;
            int gv = g.grid()[(ay + 1) * g.P + ax + 1];
            if (gv == 0 || gv == 0xFF) return false;
            return enemy_in_range(g, e.w, g.w0()[gv - 1], e.range * e.range);
        }
    }
    return false;
}

// sum of in-flight PRODUCE costs per player == resourcesUsed of GameState.getResourceUsage (GameState.java:652-664)
DEV void reserved_resources(const Game &g, int &r0, int &r1) {
    int n = g.hdr()[H_NUNITS], a = 0, b = 0;
    #pragma unroll 1
    for (int i = g.lane; i < n; i += 32) {
        uint32_t A0 = g.a0()[i];
        if (a_type(A0) == ACT_PRODUCE) { int c = ut_cost(g, a_utype(A0)), pl = u_pl(g.w0()[i]); if (pl == 1) a += c; else if (pl == 2) b += c; } // pl 0: hidden by po_hide
    }
    r0 = __reduce_add_sync(FULLM, a); r1 = __reduce_add_sync(FULLM, b);
}

// ResourceUsage.consistentWith, resource half (ResourceUsage.java:40-47): self = the candidate action {cost for player pl},
// other = the accumulated usage par[2]
DEV bool res_consistent_cand_vs_acc(const Game &g, int pl, int cost, int par0, int par1) {
    int s0 = (pl == 1 ? cost : 0) + par0, s1 = (pl == 2 ? cost : 0) + par1;
    if (par0 != 0 && s0 > 0 && s0 > g.hdr()[H_RES0]) return false;
    if (par1 != 0 && s1 > 0 && s1 > g.hdr()[H_RES1]) return false;
    return true;
}

// Sequentially accept/reject the choices of up to 32 lanes in lane order against the accumulated PlayerAction usage
// (RandomBiasedAI.java:92-99 / PlayerAction.fromVectorAction PlayerAction.java:407-411).  Returns this lane's verdict.
DEV bool accept_in_order(Game &g, int player, int cnt, int tcell, int cost, bool candidate, int &par0, int &par1) {
    int pl = player + 1;
    // Parallel path: no candidate of this chunk costs resources.  Then `par` does not change, the resource half of
    // consistentWith is the same for every lane, and verdicts only interact when two lanes want the same cell (the
    // first in lane order wins; if the first is refused because the cell is taken, so are the others).
    if (__ballot_sync(FULLM, candidate && cost != 0) == 0) {
        bool over = (par0 > 0 && par0 > g.hdr()[H_RES0]) || (par1 > 0 && par1 > g.hdr()[H_RES1]);
        bool ok = candidate && !over;
        int key = (ok && tcell >= 0) ? tcell : -1 - g.lane;
        unsigned same = __match_any_sync(FULLM, key);
        if (ok && tcell >= 0) {
            if (g.resv()[tcell] != 0 || ((g.claim()[tcell] >> player) & 1)) ok = false;
            else if (same & ((1u << g.lane) - 1)) ok = false;
        }
        __syncwarp(); // all lanes have read claim[] before it is updated
        if (ok && tcell >= 0) g.claim()[tcell] |= (uint8_t)(1 << player);
        __syncwarp();
        return ok;
    }
    bool mine = false;
    #pragma unroll 1
    for (int j = 0; j < cnt; j++) {
        int c = __shfl_sync(FULLM, tcell, j), co = __shfl_sync(FULLM, cost, j);
        bool cand = __shfl_sync(FULLM, candidate ? 1 : 0, j) != 0;
        bool ok = cand;
        if (ok && c >= 0 && (g.resv()[c] != 0 || ((g.claim()[c] >> player) & 1))) ok = false;
        if (ok && !res_consistent_cand_vs_acc(g, pl, co, par0, par1)) ok = false;
        __syncwarp(); // every lane has read claim[] before lane 0 updates it
        if (ok) {
            if (pl == 1) par0 += co; else par1 += co;
            if (c >= 0 && g.lane == 0) g.claim()[c] |= (uint8_t)(1 << player);
        }
        if (g.lane == j) mine = ok;
        __syncwarp();
    }
    return mine;
}

// ---- RandomBiasedAI.getAction (ai/RandomBiasedAI.java:51-107) ------------------------------------------------------------
// Appends (unit, action) pairs for `player` to the pending list starting at pn; returns the new count.
DEVN int policy_random_biased(Game &g, int player, int pn) {
    int n = g.hdr()[H_NUNITS], n_idle = 0;
    #pragma unroll 1
    for (int base = 0; base < n; base += 32) {
        int i = base + g.lane;
        bool idle = i < n && u_pl(g.w0()[i]) == player + 1 && a_type(g.a0()[i]) == AT_IDLE;
        unsigned m = __ballot_sync(FULLM, idle);
        if (idle) g.list()[n_idle + __popc(m & ((1u << g.lane) - 1))] = (uint8_t)i;
        n_idle += __popc(m);
    }
    __syncwarp();
    if (n_idle == 0) return pn; // canExecuteAnyAction false: no RNG draw
    int par0, par1;
    reserved_resources(g, par0, par1);
    uint64_t s0 = hdr_rng(g, H_RNGP_LO);
    #pragma unroll 1
    for (int kb = 0; kb < n_idle; kb += 32) {
        int k = kb + g.lane;
        bool active = k < n_idle;
        uint32_t A0 = ACT_NONE | A0_NOUT; int A1 = 10, tcell = -1, cost = 0, s = 0;
        if (active) {
            s = g.list()[k];
            Enum e; enumerate(g, s, e);
            uint64_t st = lcg_jump(g, s0, k);
            double dr = lcg_next_double(st);
            pick_action(g, e, sample_index(dr, e.nb, e.total), 10, A0, A1, tcell, cost);
        }
        int cnt = n_idle - kb; if (cnt > 32) cnt = 32;
        bool ok = accept_in_order(g, player, cnt, tcell, cost, active, par0, par1);
        if (active) {
            if (!ok) { A0 = ACT_NONE | A0_NOUT; A1 = 10; }
            g.pslot()[pn + k] = (uint8_t)s; g.pa0()[pn + k] = A0; g.pa1()[pn + k] = A1;
        }
    }
    if (g.lane == 0) hdr_set_rng(g, H_RNGP_LO, lcg_jump(g, s0, n_idle));
    __syncwarp();
    return pn + n_idle;
}

// ---- external actions -----------------------------------------------------------------------------------------------------
// FMT_VECTOR: PlayerAction.fromVectorAction (PlayerAction.java:384-417) + UnitAction.fromVectorAction (UnitAction.java:675-709)
//             + JNIAI padding with NONE(fill) (ai/jni/JNIAI.java:51-55).   FMT_RAW: the pairs as given.
DEVN int decode_external(Game &g, int player, int pn, const int32_t *rows, int count, int format, int fill, int maxR) {
    int cells = g.W * g.H;
    int start = pn;
    // The pending list holds `cap` entries for both players together.  A player's list without duplicate rows never exceeds its
    // unit count; rows addressing the same idle unit twice are all accepted (as PlayerAction.fromVectorAction does), so the
    // first list is cut where it would eat the room of the other player's units, the second at the capacity: rows beyond that are
    // dropped and the game is flagged GE_BAD_ACTION.
    int limit = g.cap;
    if (start == 0) {
        int n = g.hdr()[H_NUNITS], opp = 0;
        #pragma unroll 1
        for (int i = g.lane; i < n; i += 32) opp += u_pl(g.w0()[i]) == 2 - player ? 1 : 0;
        limit -= __reduce_add_sync(FULLM, opp);
    }
    if (format == FMT_RAW) {
        #pragma unroll 1
        for (int kb = 0; kb < count; kb += 32) {
            int k = kb + g.lane;
            bool ok = false; uint32_t A0 = 0; int A1 = 0, s = 0;
            if (k < count) {
                const int32_t *a = rows + (long long)k * 8;
                int cell = a[0], at = a[1];
                if (cell >= 0 && cell < cells && at >= 0 && at <= 5) {
                    int gv = g.grid()[(cell / g.W + 1) * g.P + cell % g.W + 1];
                    if (gv != 0 && gv != 0xFF) {
                        s = gv - 1; ok = true;
                        int ut = (at == ACT_PRODUCE) ? a[5] : 0xFF;
                        if (at == ACT_PRODUCE && (ut < 0 || ut >= MRTS_MAX_TYPES)) ok = false;
                        A0 = (uint32_t)at | ((uint32_t)(ut & 0xff) << 8);
                        if (at == ACT_ATTACK) A0 |= ((uint32_t)(a[3] & 0xff) << 16) | ((uint32_t)(a[4] & 0xff) << 24);
                        A1 = a[2];
                    }
                }
            }
            unsigned m = __ballot_sync(FULLM, ok);
            if (k < count && !ok) atomicOr(&g.hdr()[H_ERR], GE_BAD_ACTION);
            if (ok) { int q = pn + __popc(m & ((1u << g.lane) - 1)); if (q < limit) { g.pslot()[q] = (uint8_t)s; g.pa0()[q] = A0; g.pa1()[q] = A1; } }
            pn += __popc(m);
            if (pn > limit) { pn = limit; if (g.lane == 0) atomicOr(&g.hdr()[H_ERR], GE_BAD_ACTION); } // the pending list holds `cap` entries: rows beyond it (duplicates) are dropped and flagged
        }
        __syncwarp();
        return pn;
    }
    int par0, par1;
    reserved_resources(g, par0, par1);
    int R = maxR, ctr = R / 2;
    #pragma unroll 1
    for (int kb = 0; kb < count; kb += 32) {
        int k = kb + g.lane;
        bool cand = false; uint32_t A0 = 0; int A1 = -1, s = 0, tcell = -1, cost = 0;
        if (k < count) {
            const int32_t *a = rows + (long long)k * 8;
            int cell = a[0], at = a[1];
            if (cell >= 0 && cell < cells) {
                int gv = g.grid()[(cell / g.W + 1) * g.P + cell % g.W + 1];
                if (gv != 0 && gv != 0xFF) {
                    s = gv - 1;
                    uint32_t w = g.w0()[s];
                    if (u_pl(w) == player + 1 && a_type(g.a0()[s]) == AT_IDLE && (at < 0 || at > 5)) atomicOr(&g.hdr()[H_ERR], GE_BAD_ACTION);
                    if (u_pl(w) == player + 1 && a_type(g.a0()[s]) == AT_IDLE && at >= 0 && at <= 5) {
                        cand = true;
                        A0 = (uint32_t)at | A0_NOUT;
                        switch (at) {
                            case ACT_MOVE: A1 = a[2]; break;
                            case ACT_HARVEST: A1 = a[3]; break;
                            case ACT_RETURN: A1 = a[4]; break;
                            case ACT_PRODUCE: A1 = a[5]; A0 = ACT_PRODUCE | ((uint32_t)(a[6] & 0xff) << 8);
                                if (a[6] < 0 || a[6] >= MRTS_MAX_TYPES) cand = false; else cost = ut_cost(g, a[6]); break;
                            case ACT_ATTACK: {
                                int ax = u_x(w) + (a[7] % R - ctr), ay = u_y(w) + (a[7] / R - ctr);
                                A0 |= ((uint32_t)(ax & 0xff) << 16) | ((uint32_t)(ay & 0xff) << 24);
                                if (ax < 0 || ay < 0 || ax > 255 || ay > 255) A0 = ACT_ATTACK | A0_NOUT | (0xFFu << 16) | (0xFFu << 24);
                            } break;
                        }
                        if (a_uses_cell(at)) tcell = linear_target_cell(g, w, A1);
                        if (!cand) atomicOr(&g.hdr()[H_ERR], GE_BAD_ACTION);
                    }
                }
            }
        }
        int cnt = count - kb; if (cnt > 32) cnt = 32;
        bool ok = accept_in_order(g, player, cnt, tcell, cost, cand, par0, par1);
        unsigned m = __ballot_sync(FULLM, ok);
        if (ok) { int q = pn + __popc(m & ((1u << g.lane) - 1)); if (q < limit) { g.pslot()[q] = (uint8_t)s; g.pa0()[q] = A0; g.pa1()[q] = A1; } }
        pn += __popc(m);
        if (pn > limit) { pn = limit; if (g.lane == 0) atomicOr(&g.hdr()[H_ERR], GE_BAD_ACTION); }
        __syncwarp();
    }
    // drop this player's tentative claims (they only model PlayerAction.r while the action list is being built)
    #pragma unroll 1
    for (int k = g.lane; k < count; k += 32) {
        const int32_t *a = rows + (long long)k * 8;
        int cell = a[0], at = a[1];
        if (cell >= 0 && cell < cells && a_uses_cell(at)) {
            int gv = g.grid()[(cell / g.W + 1) * g.P + cell % g.W + 1];
            if (gv != 0 && gv != 0xFF) { int tc = linear_target_cell(g, g.w0()[gv - 1], at == ACT_MOVE ? a[2] : a[5]); if (tc >= 0) g.claim()[tc] = 0; }
        }
    }
    __syncwarp();
    if (fill >= 0) { // PlayerAction.fillWithNones (PlayerAction.java:217-235): idle own units not already in the action
        int n = g.hdr()[H_NUNITS];
        #pragma unroll 1
        for (int base = 0; base < n; base += 32) {
            int i = base + g.lane;
            bool idle = i < n && u_pl(g.w0()[i]) == player + 1 && a_type(g.a0()[i]) == AT_IDLE;
            if (idle) for (int q = start; q < pn; q++) if (g.pslot()[q] == i) { idle = false; break; }
            unsigned m = __ballot_sync(FULLM, idle);
            if (idle) { int q = pn + __popc(m & ((1u << g.lane) - 1)); if (q < limit) { g.pslot()[q] = (uint8_t)i; g.pa0()[q] = ACT_NONE | A0_NOUT; g.pa1()[q] = fill; } }
            pn += __popc(m);
            if (pn > limit) { pn = limit; if (g.lane == 0) atomicOr(&g.hdr()[H_ERR], GE_BAD_ACTION); }
            __syncwarp();
        }
    }
    return pn;
}

// GameState.issueSafe legality pass (GameState.java:347-354,386-399): illegal actions become NONE(ETA(action)).
DEV void legality_pass(Game &g, int from, int to) {
    #pragma unroll 1
    for (int k = from + g.lane; k < to; k += 32) {
        int s = g.pslot()[k];
        uint32_t A0 = g.pa0()[k]; int A1 = g.pa1()[k];
        if (!action_is_legal(g, s, A0, A1)) {
            g.pa1()[k] = eta_of(g, u_type(g.w0()[s]), A0, A1);
            g.pa0()[k] = ACT_NONE | A0_NOUT;
        }
    }
    __syncwarp();
}

// ---- GameState.issue (GameState.java:249-328) --------------------------------------------------------------------------
// One conflicting pair (existing assignment of slot e, new action N of a unit of type t).  Uniform across lanes.
DEV void resolve_conflict(Game &g, int e, int t, uint32_t &A0, int &A1, int time) {
    uint32_t E0 = g.a0()[e]; int E1 = g.a1()[e]; uint32_t ew = g.w0()[e];
    bool same_cycle = g.tis()[e] == time;
    bool cancel_old = false, cancel_new = false;
    if (same_cycle) {
        switch (g.conflict) {
            default: cancel_old = cancel_new = true; break;                         // CANCEL_BOTH
            case 2: {                                                               // CANCEL_RANDOM: GameState.r.nextInt(2)
                uint64_t s = hdr_rng(g, H_RNGC_LO);
                int r = lcg_next_int_bound(s, 2);
                __syncwarp();
                if (g.lane == 0) hdr_set_rng(g, H_RNGC_LO, s);
                if (r == 0) cancel_new = true; else cancel_old = true;
            } break;
            case 3: {                                                               // CANCEL_ALTERNATING
                int ctr = g.hdr()[H_CANCELCTR];
                __syncwarp();
                if (g.lane == 0) g.hdr()[H_CANCELCTR] = ctr + 1;
                if ((ctr % 2) == 0) cancel_new = true; else cancel_old = true;
            } break;
        }
    }
    int d1 = eta_of(g, u_type(ew), E0, E1), d2 = eta_of(g, t, A0, A1);
    int d = d1 < d2 ? d1 : d2;
    __syncwarp();
    if (same_cycle) {
        if (cancel_old && g.lane == 0) {
            if (a_uses_cell(a_type(E0))) { int tc = target_cell(g, cell_of(g, ew), E1); if (g.resv()[tc] == e + 1) g.resv()[tc] = 0; }
            g.a0()[e] = (E0 & 0xF0u) | ACT_NONE | A0_NOUT; g.a1()[e] = d; g.rdy()[e] = time + d;
        }
        if (cancel_new) { A0 = ACT_NONE | A0_NOUT; A1 = d; }
    } else {
        if (g.lane == 0) g.hdr()[H_ERR] |= GE_INCONSISTENT_OLDER;
        A0 = ACT_NONE | A0_NOUT; A1 = -1; // new UnitAction(TYPE_NONE): parameter stays -1 (GameState.java:316)
    }
    __syncwarp();
}

DEVN void issue_pending(Game &g, int from, int to) {
    int time = g.hdr()[H_TIME];
    #pragma unroll 1
    for (int k = from; k < to; k++) {
        int s = g.pslot()[k];
        uint32_t A0 = g.pa0()[k]; int A1 = g.pa1()[k];
        uint32_t w = g.w0()[s];
        int t = u_type(w), pl = u_pl(w), c = cell_of(g, w), at = a_type(A0);
        int tc = -1, cost = 0;
        if (a_uses_cell(at)) { tc = target_cell(g, c, A1); if (at == ACT_PRODUCE) cost = ut_cost(g, a_utype(A0)); }
        __syncwarp();
        if (tc >= 0 && g.lane == 0) g.claim()[tc] = 0;
        if (tc >= 0 && cost == 0) {
            int e = g.resv()[tc]; // at most one in-flight action can hold a cell
            if (e != 0) resolve_conflict(g, e - 1, t, A0, A1, time);
        } else if (cost != 0) {
            // general pairwise check against every existing assignment, in insertion order (GameState.java:263-319)
            int n = g.hdr()[H_NUNITS], pres = pl ? g.hdr()[H_RES0 + pl - 1] : 0;
            unsigned cb = 0; // bit j: slot j*32+lane conflicts
            #pragma unroll 1
            for (int j = 0; j * 32 < n; j++) {
                int i = j * 32 + g.lane;
                if (i < n) {
                    uint32_t E0 = g.a0()[i]; int eat = a_type(E0);
                    if (eat != AT_IDLE && !(E0 & A0_DEAD)) {
                        uint32_t ew = g.w0()[i];
                        bool pos = a_uses_cell(eat) && target_cell(g, cell_of(g, ew), g.a1()[i]) == tc;
                        int ecost = (eat == ACT_PRODUCE && u_pl(ew) == pl) ? ut_cost(g, a_utype(E0)) : 0;
                        bool res = (ecost + cost > 0) && (ecost + cost > pres);
                        if (pos || res) cb |= 1u << j;
                    }
                }
            }
            #pragma unroll 1
            for (;;) {
                uint32_t best = 0xFFFFFFFFu; int bj = 0;
                #pragma unroll 1
                for (unsigned m = cb; m; m &= m - 1) { int j = __ffs(m) - 1; uint32_t q = g.seq()[j * 32 + g.lane]; if (q < best) { best = q; bj = j; } }
                uint32_t mn = __reduce_min_sync(FULLM, best);
                if (mn == 0xFFFFFFFFu) break;
                int owner = __ffs(__ballot_sync(FULLM, best == mn)) - 1;
                int e = __shfl_sync(FULLM, bj * 32 + g.lane, owner);
                if (g.lane == owner) cb &= ~(1u << bj);
                resolve_conflict(g, e, t, A0, A1, time);
            }
        }
        // unitActions.put(unit, new UnitActionAssignment(unit, action, time)): an existing key keeps its slot
        __syncwarp(); // every lane has finished reading resv/a0 for this action
        if (g.lane == 0) {
            uint32_t prev = g.a0()[s];
            if (a_type(prev) == AT_IDLE) g.seq()[s] = (uint32_t)g.hdr()[H_NEXTSEQ]++;
            else if (a_uses_cell(a_type(prev))) { int ptc = target_cell(g, c, g.a1()[s]); if (g.resv()[ptc] == s + 1) g.resv()[ptc] = 0; }
            g.a0()[s] = (prev & 0xF0u) | (A0 & ~0xF0u); g.a1()[s] = A1; g.tis()[s] = time; g.rdy()[s] = time + eta_of(g, t, A0, A1);
            if (a_uses_cell(a_type(A0))) g.resv()[tc] = (uint8_t)(s + 1);
        }
        __syncwarp();
    }
}

// Parallel issue of the two action lists produced by device policies in one cycle, valid under CANCEL_BOTH.
// Why it is exact (GameState.issue, GameState.java:249-328): a policy only emits actions that are consistent with the
// in-flight usage and with its own earlier choices (RandomBiasedAI.java:92-99), and both policies looked at the same
// pre-issue state (Game.java:134-137).  So the only inconsistencies issue() can find are between a player-1 action and a
// player-0 action of this same cycle that target the same cell -- isolated pairs, each cancelled to
// NONE(min(ETA_old, ETA_new)) -- and insertion order is simply list order.
DEVN void issue_policy_lists(Game &g, int pn0, int pn1) {
    int time = g.hdr()[H_TIME];
    uint32_t base_seq = (uint32_t)g.hdr()[H_NEXTSEQ];
    #pragma unroll 1
    for (int k = g.lane; k < pn0; k += 32) {
        int s = g.pslot()[k];
        uint32_t A0 = g.pa0()[k]; int A1 = g.pa1()[k];
        if (a_uses_cell(a_type(A0))) { int tc = target_cell(g, cell_of(g, g.w0()[s]), A1); g.resv()[tc] = (uint8_t)(s + 1); g.claim()[tc] = 0; }
        g.a0()[s] = (g.a0()[s] & 0xF0u) | (A0 & ~0xF0u); g.a1()[s] = A1; g.tis()[s] = time; g.seq()[s] = base_seq + k;
        g.rdy()[s] = time + eta_of(g, u_type(g.w0()[s]), A0, A1);
    }
    __syncwarp();
    #pragma unroll 1
    for (int k = pn0 + g.lane; k < pn1; k += 32) {
        int s = g.pslot()[k];
        uint32_t A0 = g.pa0()[k]; int A1 = g.pa1()[k];
        uint32_t w = g.w0()[s];
        if (a_uses_cell(a_type(A0))) {
            int tc = target_cell(g, cell_of(g, w), A1);
            g.claim()[tc] = 0;
            int e = g.resv()[tc];
            if (e == 0) g.resv()[tc] = (uint8_t)(s + 1);
            else {
                e--;
                if (g.tis()[e] == time) {
                    uint32_t E0 = g.a0()[e];
                    int d1 = eta_of(g, u_type(g.w0()[e]), E0, g.a1()[e]), d2 = eta_of(g, u_type(w), A0, A1);
                    int d = d1 < d2 ? d1 : d2;
                    g.a0()[e] = (E0 & 0xF0u) | ACT_NONE | A0_NOUT; g.a1()[e] = d; g.rdy()[e] = time + d;
                    g.resv()[tc] = 0;
                    A0 = ACT_NONE | A0_NOUT; A1 = d;
                } else {
                    atomicOr(&g.hdr()[H_ERR], GE_INCONSISTENT_OLDER);
                    A0 = ACT_NONE | A0_NOUT; A1 = -1;
                }
            }
        }
        g.a0()[s] = (g.a0()[s] & 0xF0u) | (A0 & ~0xF0u); g.a1()[s] = A1; g.tis()[s] = time; g.seq()[s] = base_seq + k;
        g.rdy()[s] = time + eta_of(g, u_type(w), A0, A1);
    }
    __syncwarp();
    if (g.lane == 0) g.hdr()[H_NEXTSEQ] = (int32_t)(base_seq + pn1);
    __syncwarp();
}

// ---- GameState.cycle (GameState.java:553-571) + UnitAction.execute (UnitAction.java:338-465) ----------------------------
DEV void kill_unit(Game &g, int v) { // GameState.removeUnit (GameState.java:79-82); lane 0 only
    uint32_t vw = g.w0()[v]; uint32_t V0 = g.a0()[v];
    int vc = cell_of(g, vw);
    g.grid()[vc] = 0; if (!g.slim) g.kind()[vc] = 0;
    if (a_uses_cell(a_type(V0))) { int tc = target_cell(g, vc, g.a1()[v]); if (g.resv()[tc] == v + 1) g.resv()[tc] = 0; }
    g.a0()[v] = V0 | A0_DEAD; // keeps its action words: a ready action of a dead unit still executes this cycle
}

DEV int neighbour_slot(const Game &g, int c, int dir) { // getUnitAt of the adjacent cell, -1 if none
    if ((unsigned)dir >= 4u) return -1;
    int gv = g.grid()[c + doff(g, dir)];
    return (gv == 0 || gv == 0xFF) ? -1 : gv - 1;
}

// Execute the (already removed) assignment (A0,A1) of slot s.  Called by ONE lane at a time, for every ready assignment in
// insertion order, so it is plain sequential code with no warp primitives.
DEV void execute_serial(Game &g, int s, int &ndead) {
    uint32_t A0 = g.a0()[s]; int A1 = g.a1()[s];
    uint32_t w = g.w0()[s];
    bool dead = (A0 & A0_DEAD) != 0;
    int c = cell_of(g, w);
    // unitActions.remove(uaa.unit) (GameState.java:563); a dead unit lost its entry (and reservation) when it died
    g.a0()[s] = (A0 & 0xF0u) | AT_IDLE | A0_NOUT; g.rdy()[s] = MRTS_NEVER;
    if (a_type(A0) == ACT_MOVE && !dead && (unsigned)A1 < 4u) {
        // the common case on its own short path: a live unit steps into the cell it had reserved (UnitAction.java:346-361)
        int nc = c + doff(g, A1);
        int rv = g.resv()[nc], gv = g.grid()[nc], kv = g.slim ? 0 : g.kind()[c]; // three independent loads, issued together
        if (rv == s + 1) g.resv()[nc] = 0;
        if (gv != 0) g.hdr()[H_ERR] |= GE_CELL_OCCUPIED;
        else {
            g.grid()[c] = 0; g.grid()[nc] = (uint8_t)(s + 1);
            if (!g.slim) { g.kind()[nc] = (uint8_t)kv; g.kind()[c] = 0; }
            // x +- 1 or y +- 1 inside the packed word: a legal move never leaves [0, 255], so no carry crosses a field
            g.w0()[s] = w + ((A1 & 1) ? (uint32_t)(2 - A1) << 16 : (uint32_t)(A1 - 1) << 24);
        }
        return;
    }
    int t = u_type(w), pl = u_pl(w);
    if (!dead && a_uses_cell(a_type(A0))) { int tc = target_cell(g, c, A1); if (g.resv()[tc] == s + 1) g.resv()[tc] = 0; }
    switch (a_type(A0)) {
        case ACT_MOVE:
            if (!dead && (unsigned)A1 < 4u) {
                int nc = c + doff(g, A1);
                if (g.grid()[nc] != 0) g.hdr()[H_ERR] |= GE_CELL_OCCUPIED;
                else {
                    g.grid()[c] = 0; g.grid()[nc] = (uint8_t)(s + 1);
                    if (!g.slim) { g.kind()[nc] = g.kind()[c]; g.kind()[c] = 0; }
                    g.w0()[s] = (w & 0xffffu) | ((uint32_t)(u_x(w) + ddx(A1)) << 16) | ((uint32_t)(u_y(w) + ddy(A1)) << 24);
                }
            }
            break;
        case ACT_ATTACK: {
            int ax = (A0 >> 16) & 0xff, ay = A0 >> 24;
            if (ax < g.W && ay < g.H) {
                int gv = g.grid()[(ay + 1) * g.P + ax + 1];
                if (gv != 0 && gv != 0xFF) {
                    int v = gv - 1;
                    int mn = ut_mind(g, t), mx = ut_maxd(g, t), dmg = mn;
                    if (mn != mx) { // UnitAction.r.nextInt(1 + max - min)
                        uint64_t rs = hdr_rng(g, H_RNGD_LO);
                        dmg = mn + lcg_next_int_bound(rs, 1 + (mx - mn));
                        hdr_set_rng(g, H_RNGD_LO, rs);
                    }
                    uint32_t vw1 = g.w1()[v];
                    int hp = u_hp(vw1) - dmg;
                    g.w1()[v] = mk_w1(hp, u_res(vw1));
                    if (hp <= 0) { kill_unit(g, v); ndead++; }
                }
            }
        } break;
        case ACT_HARVEST: {
            int r = neighbour_slot(g, c, A1);
            if (r >= 0) {
                uint32_t rw1 = g.w1()[r], mw1 = g.w1()[s];
                if ((ut_flags(g, u_type(g.w0()[r])) & UF_RESOURCE) && (ut_flags(g, t) & UF_HARVEST) && u_res(mw1) == 0) {
                    int amt = ut_hamt(g, t), left = u_res(rw1) - amt;
                    g.w1()[r] = mk_w1(u_hp(rw1), left);
                    if (left <= 0) { kill_unit(g, r); ndead++; }
                    g.w1()[s] = mk_w1(u_hp(mw1), amt);
                }
            }
        } break;
        case ACT_RETURN: {
            int b = neighbour_slot(g, c, A1);
            if (b >= 0 && pl != 0) {
                uint32_t mw1 = g.w1()[s];
                if ((ut_flags(g, u_type(g.w0()[b])) & UF_STOCKPILE) && u_res(mw1) > 0) {
                    g.hdr()[H_RES0 + pl - 1] += u_res(mw1);
                    g.w1()[s] = mk_w1(u_hp(mw1), 0);
                }
            }
        } break;
        case ACT_PRODUCE: {
            int ut = a_utype(A0);
            if (pl != 0 && ut < MRTS_MAX_TYPES) {
                int n = g.hdr()[H_NUNITS], pres = g.hdr()[H_RES0 + pl - 1];
                int id = g.hdr()[H_NEXTID]++; // new Unit(...) takes an ID even when the unit is then not added
                int cost = ut_cost(g, ut);
                if (pres - cost >= 0) {
                    int nc = target_cell(g, c, A1);
                    if ((unsigned)A1 >= 4u || g.grid()[nc] != 0) g.hdr()[H_ERR] |= GE_CELL_OCCUPIED;
                    else if (n >= g.cap) g.hdr()[H_ERR] |= GE_UNIT_OVERFLOW;
                    else {
                        int nx = u_x(w) + ddx(A1), ny = u_y(w) + ddy(A1);
                        g.w0()[n] = (uint32_t)ut | ((uint32_t)pl << 8) | ((uint32_t)nx << 16) | ((uint32_t)ny << 24);
                        g.w1()[n] = mk_w1(ut_hp(g, ut), 0);
                        g.a0()[n] = AT_IDLE | A0_NOUT; g.a1()[n] = 0; g.tis()[n] = 0; g.seq()[n] = 0; g.uid()[n] = (uint32_t)id;
                        if (g.uw > MRTS_UNIT_WORDS_CORE) { g.x0()[n] = 0; g.x1()[n] = 0; }
                        g.grid()[nc] = (uint8_t)(n + 1); if (!g.slim) g.kind()[nc] = (uint8_t)kind_of(g, g.w0()[n]); g.rdy()[n] = MRTS_NEVER;
                        g.hdr()[H_NUNITS] = n + 1;
                        g.hdr()[H_RES0 + pl - 1] = pres - cost;
                    }
                } else g.hdr()[H_ERR] |= GE_FAILED_PRODUCE;
            }
        } break;
        default: break;
    }
}

// remove dead slots, keeping order; rebuild the cell maps
DEV void compact_units(Game &g) {
    int n = g.hdr()[H_NUNITS], out = 0;
    uint32_t *su = g.w0();
    #pragma unroll 1
    for (int base = 0; base < n; base += 32) {
        int i = base + g.lane;
        bool alive = i < n && !(g.a0()[i] & A0_DEAD);
        uint32_t r[MRTS_UNIT_WORDS + 1]; // the words mirrored in HBM + RDY
        if (alive) for (int k = 0; k <= MRTS_UNIT_WORDS; k++) if (k <= g.uw) r[k] = su[k * g.cap + i];
        unsigned m = __ballot_sync(FULLM, alive);
        int pos = out + __popc(m & ((1u << g.lane) - 1));
        if (i < n) g.list()[i] = alive ? (uint8_t)(pos + 1) : 0; // old slot -> new slot + 1 (0: removed)
        __syncwarp();
        if (alive) for (int k = 0; k <= MRTS_UNIT_WORDS; k++) if (k <= g.uw) su[k * g.cap + pos] = r[k];
        out += __popc(m);
        __syncwarp();
    }
    if (g.lane == 0) g.hdr()[H_NUNITS] = out;
    __syncwarp();
    // unit references held by abstract actions (X1: attack/harvest target, base) follow their unit or become "dead object"
#pragma unroll 1
    for (int i = g.lane; i < (g.uw > MRTS_UNIT_WORDS_CORE ? out : 0); i += 32) {
        uint32_t X1 = g.x1()[i];
        uint32_t t = X1 & 0xff, b = (X1 >> 8) & 0xff;
        if (t != 0 && t != 0xFF) { t = g.list()[t - 1]; if (t == 0) t = 0xFF; }
#ifndef MRTS_TU_RUSH_ONLY
        if ((g.x0()[i] & 7u) == 7u) { g.x1()[i] = (X1 & 0xffffff00u) | t; continue; } // AA_TACTIC (scripted.cuh): the base field is a coordinate
#endif
        if (b != 0 && b != 0xFF) { b = g.list()[b - 1]; if (b == 0) b = 0xFF; }
        g.x1()[i] = (X1 & 0xffff0000u) | (b << 8) | t;
    }
    g_rebuild(g, false);
}

// PhysicalGameState.gameover / winner (PhysicalGameState.java:334-387); returns gameover, sets winner (-1 none)
DEV bool game_over(const Game &g, int &winner) {
    int n = g.hdr()[H_NUNITS], c0 = 0, c1 = 0;
    #pragma unroll 1
    for (int base = 0; base < n; base += 32) {
        int i = base + g.lane;
        int pl = i < n ? u_pl(g.w0()[i]) : 0;
        c0 += __popc(__ballot_sync(FULLM, pl == 1));
        c1 += __popc(__ballot_sync(FULLM, pl == 2));
    }
    winner = (c0 > 0 && c1 == 0) ? 0 : ((c1 > 0 && c0 == 0) ? 1 : -1);
    return c0 == 0 || c1 == 0;
}

// earliest completion time of any in-flight assignment (GameState.getNextChangeTime, GameState.java:539-542)
DEV int warp_min_time(int best) { return (int)(__reduce_min_sync(FULLM, (unsigned)(best ^ 0x80000000)) ^ 0x80000000u); }
DEV int min_ready_time(const Game &g) {
    int n = g.hdr()[H_NUNITS], best = MRTS_NEVER;
    #pragma unroll 1
    for (int i = g.lane; i < n; i += 32) { int r = g.rdy()[i]; if (r < best) best = r; }
    return warp_min_time(best);
}

// time := t_new, then execute every assignment with ETA + issueTime <= time in insertion order (GameState.cycle).
// Returns non-zero when units were removed; the caller re-evaluates gameover() only then (nothing else can end a game).
// Ready assignments are found in parallel; NONE actions (no effect, so their position in the order is irrelevant) are
// retired on the spot; the rest executes one at a time in insertion-sequence order (a warp-wide min picks the next
// one), because the reference's effects are order dependent (kills, depletion, produce/return on a player's resources).
DEV int cycle_execute(Game &g, int t_new) {
    __syncwarp();
    if (g.lane == 0) g.hdr()[H_TIME] = t_new;
    int n = g.hdr()[H_NUNITS];
    // this lane's ready assignment with the smallest insertion sequence (a lane owns units lane, lane + 32, ...)
    uint32_t cs = 0xFFFFFFFFu; int ci = 0, mine = 0; // mine: ready assignments this lane still has to execute
    #pragma unroll 1
    for (int i = g.lane; i < n; i += 32) {
        if (g.rdy()[i] <= t_new) {
            uint32_t A0 = g.a0()[i];
            if (a_type(A0) == ACT_NONE) { g.a0()[i] = (A0 & 0xF0u) | AT_IDLE | A0_NOUT; g.rdy()[i] = MRTS_NEVER; }
            else { uint32_t q = g.seq()[i]; mine++; if (q < cs) { cs = q; ci = i; } }
        }
    }
    // The ready assignments execute in insertion-sequence order, but a run of well-formed MOVEs -- a live unit stepping
    // into the empty cell that it alone has reserved -- touches disjoint cells and nothing else, so the MOVEs of a run
    // commute.  Each round: every owner lane re-examines its candidate on the current state; all well-formed MOVEs older
    // than the oldest other assignment execute at once, then that assignment executes alone (it may kill, deplete, produce
    // or be a MOVE that is not well-formed).  A cycle of MOVEs only -- the common case -- takes one round.
    int ndead = 0;
    #pragma unroll 1
    for (;;) {
        bool is_par = false;
        uint32_t w = 0, A0 = 0; int A1 = 0, c = 0, nc = 0, kv = 0;
        if (cs != 0xFFFFFFFFu && mine == 1) {
            A0 = g.a0()[ci];
            if (a_type(A0) == ACT_MOVE && !(A0 & A0_DEAD)) {
                A1 = g.a1()[ci]; w = g.w0()[ci];
                if ((unsigned)A1 < 4u) {
                    c = cell_of(g, w); nc = c + doff(g, A1);
                    int rv = g.resv()[nc], gv = g.grid()[nc]; kv = g.slim ? 0 : g.kind()[c];
                    is_par = rv == ci + 1 && gv == 0;
                }
            }
        }
        uint32_t s_np = __reduce_min_sync(FULLM, is_par ? 0xFFFFFFFFu : cs); // oldest assignment that must execute alone
        if (is_par && cs < s_np) { // (the reduction also orders every lane's loads before these stores)
            g.a0()[ci] = (A0 & 0xF0u) | AT_IDLE | A0_NOUT; g.rdy()[ci] = MRTS_NEVER;
            g.resv()[nc] = 0;
            g.grid()[c] = 0; g.grid()[nc] = (uint8_t)(ci + 1);
            if (!g.slim) { g.kind()[nc] = (uint8_t)kv; g.kind()[c] = 0; }
            g.w0()[ci] = w + ((A1 & 1) ? (uint32_t)(2 - A1) << 16 : (uint32_t)(A1 - 1) << 24);
            cs = 0xFFFFFFFFu; mine = 0;
        }
        __syncwarp();
        if (s_np == 0xFFFFFFFFu) break; // nothing had to execute alone: the round took everything that was left
        if (cs == s_np) { // the owner of the oldest remaining assignment executes it, then looks for its next one
            execute_serial(g, ci, ndead);
            cs = 0xFFFFFFFFu;
            if (--mine > 0) { // rare: this lane owns another ready unit (units lane, lane + 32, ...)
                #pragma unroll 1
                for (int i = g.lane; i < n; i += 32)
                    if (g.rdy()[i] <= t_new) { uint32_t q = g.seq()[i]; if (q < cs) { cs = q; ci = i; } }
            }
        }
        __syncwarp(); // its effects are visible to the lanes of the next round
    }
    ndead = __ballot_sync(FULLM, ndead > 0) ? 1 : 0;
    if (ndead) compact_units(g);
    return ndead;
}
DEVN int cycle_execute_ni(Game &g, int t_new) { return cycle_execute(g, t_new); } // one out-of-line copy for the generic kernel

// ---- RandomBiasedAI.getAction for the players of one decision point, fused with issue (fast path) ----------------------
// Lane i works on unit i (no list compaction, no pending list): a player's idle units enumerate, sample and arbitrate in
// unit-list order, and the accepted actions go straight into the assignment words.
struct RbCtx {
    int time;          // GameState.time of the decision
    uint32_t seq_base; // first insertion sequence number of this decision point
    uint64_t s0;       // Sampler.generator state before the first draw
    int k;             // draws (= actions issued) so far: player 0's idle units in list order, then player 1's
    int par0, par1;    // PlayerAction.r resource usage while the current player's action is being built
    int minr;          // lane-local min completion time of the assignments written by this lane
    bool cancelled;    // lane-local: this lane's action was cancelled against the other player's (claim map needs clearing)
};

// Scan the unit table once: the idle units of each player compacted in list order (player 0's slots into list(), player
// 1's into the pslot area; cnt0 / cnt1 of them), the in-flight PRODUCE costs per player (GameState.getResourceUsage,
// GameState.java:652-664) and the lane-local minimum completion time of in-flight actions.  Returns bit 0 / bit 1: player
// 0 / 1 has an idle unit.
DEV int rb_scan(const Game &g, int n, int &par0, int &par1, int &mr, int &cnt0, int &cnt1) {
    int a = 0, b = 0;
    uint8_t *list0 = g.list(), *list1 = g.base() + g.o_pslot;
    unsigned below = (1u << g.lane) - 1;
    cnt0 = cnt1 = 0;
    mr = MRTS_NEVER;
    #pragma unroll 1
    for (int base = 0; base < n; base += 32) {
        int i = base + g.lane;
        bool i0 = false, i1 = false;
        if (i < n) {
            uint32_t A0 = g.a0()[i];
            int at = a_type(A0), pl = u_pl(g.w0()[i]), r = g.rdy()[i];
            if (at == (int)AT_IDLE) { i0 = pl == 1; i1 = pl == 2; }
            else {
                if (r < mr) mr = r;
                if (at == ACT_PRODUCE) { int c = ut_cost(g, a_utype(A0)); if (pl == 1) a += c; else b += c; }
            }
        }
        unsigned m0 = __ballot_sync(FULLM, i0), m1 = __ballot_sync(FULLM, i1);
        if (i0) list0[cnt0 + __popc(m0 & below)] = (uint8_t)i;
        if (i1) list1[cnt1 + __popc(m1 & below)] = (uint8_t)i;
        cnt0 += __popc(m0); cnt1 += __popc(m1);
    }
    int idle = (cnt0 ? 1 : 0) | (cnt1 ? 2 : 0);
    if (idle) { par0 = __reduce_add_sync(FULLM, a); par1 = __reduce_add_sync(FULLM, b); }
    __syncwarp();
    return idle;
}

// Verdicts of one chunk of candidates (lanes in `m`, unit-list order) against the action being built: the chosen cell must
// not be used (ResourceUsage.consistentWith, ResourceUsage.java:31-50) by an in-flight assignment or by an earlier choice
// of this player, and the resources must suffice (RandomBiasedAI.java:92-99).  `simul`: a cell reserved by the OTHER
// player's action of this same cycle does not block -- that player's list was built on the same state (Game.java:134-137)
// and GameState.issue will find the pair inconsistent (handled by the caller).
DEV bool rb_accept(const Game &g, int pl, unsigned m, bool cand, int tcell, int cost, bool simul, RbCtx &c) {
    // lanes that want the same cell.  MATCH.ANY has a long latency: it is issued first so that the shared-memory loads below run
    // under it (skipping it when fewer than two lanes have a cell was measured slower)
    unsigned same = __match_any_sync(FULLM, (cand && tcell >= 0) ? tcell : -1 - g.lane);
    bool blocked = false;
    if (cand && tcell >= 0) {
        int ev = g.resv()[tcell];
        if (g.slim && ev > MRTS_MAX_CAP) blocked = true; // (slim layout) the claim of a cancelled pair, coded into resv[]: see rb_player
        else if (ev != 0) blocked = !(simul && g.tis()[ev - 1] == c.time && u_pl(g.w0()[ev - 1]) != pl);
        if (!g.slim && (g.claim()[tcell] & pl)) blocked = true; // this player's earlier choice of the cell was cancelled at issue (see rb_player)
    }
    unsigned below = (1u << g.lane) - 1;
    // While no candidate that costs resources is involved, the resource half of consistentWith is the same for every lane
    // (`over`), and verdicts only interact when two lanes want the same cell: the first in list order wins, and if it is
    // blocked or over, so are the others.
    int res0 = g.hdr()[H_RES0], res1 = g.hdr()[H_RES1];
    bool over = (c.par0 > 0 && c.par0 > res0) || (c.par1 > 0 && c.par1 > res1);
    unsigned costly = __ballot_sync(FULLM, cand && cost != 0);
    if (costly == 0) return cand && !over && !blocked && !(tcell >= 0 && (same & below));
    // Candidates that cost resources change the accumulated usage, so the span from the first to the last of them is
    // decided one candidate at a time; the lanes before the span see the initial usage, the lanes after it the final one,
    // and both are decided in parallel as above.
    int firstc = __ffs(costly) - 1, lastc = 31 - __clz(costly);
    bool mine = cand && g.lane < firstc && !over && !blocked && !(tcell >= 0 && (same & below));
    unsigned acc = __ballot_sync(FULLM, mine); // accepted lanes so far
    unsigned span = m & (0xFFFFFFFFu << firstc) & (lastc == 31 ? 0xFFFFFFFFu : ((2u << lastc) - 1));
    #pragma unroll 1
    for (unsigned mm = span; mm; mm &= mm - 1) {
        int j = __ffs(mm) - 1;
        int co = __shfl_sync(FULLM, cost, j);
        unsigned sj = __shfl_sync(FULLM, same, j);
        bool ok = __shfl_sync(FULLM, blocked ? 0 : 1, j) != 0;
        if (ok && (sj & acc & ~(1u << j))) ok = false;
        if (ok && !res_consistent_cand_vs_acc(g, pl, co, c.par0, c.par1)) ok = false;
        if (ok) { if (pl == 1) c.par0 += co; else c.par1 += co; acc |= 1u << j; }
        if (g.lane == j) mine = ok;
    }
    if (g.lane > lastc) {
        over = (c.par0 > 0 && c.par0 > res0) || (c.par1 > 0 && c.par1 > res1);
        unsigned after = below & ~((2u << lastc) - 1); // earlier lanes behind the span
        mine = cand && !over && !blocked && !(tcell >= 0 && ((same & acc) || (same & after)));
    }
    return mine;
}

// One player's getAction + issue.  pl = owner code (1 = player 0, 2 = player 1); its idle units are the `cnt` slots of
// `list` (unit-list order), 32 per pass.
DEV void rb_player(Game &g, int pl, const uint8_t *list, int cnt, bool simul, RbCtx &c) {
    #pragma unroll 1
    for (int kb = 0; kb < cnt; kb += 32) {
        int left = cnt - kb;
        unsigned m = left >= 32 ? 0xFFFFFFFFu : ((1u << left) - 1);
        bool idle = (m >> g.lane) & 1;
        int i = idle ? list[kb + g.lane] : 0;
        int k = c.k + g.lane;
        uint32_t A0 = ACT_NONE | A0_NOUT; int A1 = 10, tcell = -1, cost = 0, t = 0;
        if (idle) {
            Enum e; enumerate(g, i, e);
            t = e.t;
            uint64_t st = lcg_jump(g, c.s0, k);
            double dr = lcg_next_double(st);
            pick_action(g, e, sample_index(dr, e.nb, e.total), 10, A0, A1, tcell, cost);
        }
        bool ok = rb_accept(g, pl, m, idle, tcell, cost, simul, c);
        __syncwarp(); // every lane has read resv[] before the accepted actions are written
        if (idle) {
            if (!ok) { A0 = ACT_NONE | A0_NOUT; A1 = 10; tcell = -1; }
            int eta = eta_of(g, t, A0, A1);
            if (tcell >= 0) {
                int ev = g.resv()[tcell];
                if (ev != 0) {
                    // the other player's action of this same cycle targets the cell: GameState.issue cancels both to
                    // NONE(min(ETA_old, ETA_new)) under CANCEL_BOTH (GameState.java:263-296)
                    int es = ev - 1;
                    uint32_t E0 = g.a0()[es];
                    int d1 = eta_of(g, u_type(g.w0()[es]), E0, g.a1()[es]);
                    if (d1 < eta) eta = d1;
                    g.a0()[es] = (E0 & 0xF0u) | ACT_NONE | A0_NOUT; g.a1()[es] = eta; g.rdy()[es] = c.time + eta;
                    g.resv()[tcell] = 0;
                    // the cell stays part of this player's PlayerAction.r: a later unit of the same list must not choose it
                    // (only the second player's pass can cancel, so in the slim layout one code in resv[] is enough: no slot is
                    // numbered above MRTS_MAX_CAP)
                    if (g.slim) g.resv()[tcell] = (uint8_t)(MRTS_MAX_CAP + 1); else g.claim()[tcell] |= (uint8_t)pl;
                    c.cancelled = true;
                    A0 = ACT_NONE | A0_NOUT; A1 = eta;
                } else g.resv()[tcell] = (uint8_t)(i + 1);
            }
            g.a0()[i] = (g.a0()[i] & 0xF0u) | (A0 & ~0xF0u); g.a1()[i] = A1; g.tis()[i] = c.time; g.seq()[i] = c.seq_base + k;
            g.rdy()[i] = c.time + eta;
            if (c.time + eta < c.minr) c.minr = c.time + eta;
        }
        c.k += __popc(m);
        __syncwarp();
    }
}

#include "scripted.cuh"

// ---- loops ---------------------------------------------------------------------------------------------------------------
struct WarpStats { unsigned long long v[N_WARP_STATS]; }; // lives in shared memory; only lane 0 updates it
DEV void stat_add(WarpStats &ws, int lane, int k, unsigned long long d) { if (lane == 0) ws.v[k] += d; }

// ---- PartiallyObservableGameState(gs, player) as seen by a device policy (PartiallyObservableGameState.java:35-71) --------
// MRTS_FLAG_PO_POLICIES batches (Game with partiallyObservable = true, Game.java:129-134).  Instead of building a filtered
// copy, the units the player cannot see are hidden in place for the duration of its getAction: their W0 word is parked in
// hid[] and replaced by a neutral unit of the unused type 7 (no flags, so no predicate of any policy matches it), and their
// cell and the cell their in-flight MOVE/PRODUCE reserves are cleared from the grid / kind / resv maps (removeUnit also drops
// the unit's assignment from the view).  vis[] keeps the player's sight map for the PO rushes' exploration.  Slots do not
// move, so the policy's list and abstract actions refer to the real state; po_unhide restores everything.
DEV void stamp_sight(Game &g, uint8_t *map, uint32_t w);
DEVN void po_hide(Game &g, int observer) {
    int n = g.hdr()[H_NUNITS];
    __syncwarp();
#pragma unroll 1
    for (int i = g.lane; i < g.pcw; i += 32) ((uint32_t *)g.vis())[i] = 0;
    __syncwarp();
#pragma unroll 1
    for (int i = 0; i < n; i++) { uint32_t w = g.w0()[i]; if (u_pl(w) == observer + 1) stamp_sight(g, g.vis(), w); }
    __syncwarp();
#pragma unroll 1
    for (int i = g.lane; i < n; i += 32) {
        uint32_t w = g.w0()[i];
        int c = cell_of(g, w);
        bool hide = u_pl(w) != observer + 1 && !g.vis()[c];
        g.hid()[i] = hide ? (w | 0x8000u) : 0u; // bit 15 marks the entry: a resource at (0, 0) is the all-zero word
        if (hide) {
            g.w0()[i] = (uint32_t)HIDDEN_TYPE | (w & 0xffff0000u);
            g.grid()[c] = 0; g.kind()[c] = 0;
            uint32_t A0 = g.a0()[i];
            if (a_uses_cell(a_type(A0))) { int tc = target_cell(g, c, g.a1()[i]); if (g.resv()[tc] == (uint8_t)(i + 1)) g.resv()[tc] = 0; }
        }
    }
    g.po_view = true;
    __syncwarp();
}
DEVN void po_unhide(Game &g) {
    int n = g.hdr()[H_NUNITS];
    __syncwarp();
#pragma unroll 1
    for (int i = g.lane; i < n; i += 32) {
        uint32_t w = g.hid()[i];
        if (w == 0u) continue;
        w &= ~0x8000u;
        int c = cell_of(g, w);
        g.w0()[i] = w;
        g.grid()[c] = (uint8_t)(i + 1); g.kind()[c] = (uint8_t)kind_of(g, w);
        uint32_t A0 = g.a0()[i];
        if (a_uses_cell(a_type(A0))) { int tc = target_cell(g, c, g.a1()[i]); if (g.resv()[tc] == 0) g.resv()[tc] = (uint8_t)(i + 1); }
    }
    g.po_view = false;
    __syncwarp();
}

// a device policy of a MRTS_FLAG_PO_POLICIES batch: decide on the player's view, then issueSafe on the real state (Game.java:
// 129-137): a move into a cell a hidden unit holds becomes NONE.  One out-of-line copy, away from the common path.
DEVN int run_policy_po(Game &g, const StepParams &p, int player, int pn) {
    const int pol = p.policy[player], n0 = pn;
    po_hide(g, player);
    if (pol == POL_RANDOM_BIASED) pn = policy_random_biased(g, player, pn);
    else if (p.scripted) pn = policy_scripted(g, player, pol, p.pathfinder[player], pn);
    po_unhide(g);
    legality_pass(g, n0, pn);
    return pn;
}

DEV int run_policy(Game &g, const StepParams &p, long long gi, int player, int pn, bool first_iter) {
    const int pol = p.policy[player];
#ifdef MRTS_TU_RUSH_ONLY
    { // the lean copy: both players run a scripted rush with A* (microrts_cuda.cu launches it for nothing else)
        int n0 = pn;
        pn = policy_scripted(g, player, pol, 0, pn);
        legality_pass(g, n0, pn);
        return pn;
    }
#endif
#ifndef MRTS_TU_RUSH_ONLY
    if (p.po_policies && (pol == POL_RANDOM_BIASED || POL_IS_SCRIPTED(pol))) return run_policy_po(g, p, player, pn);
#endif
    if (pol == POL_RANDOM_BIASED) return policy_random_biased(g, player, pn);
    if (pol == POL_EXTERNAL) {
        if (first_iter && p.ext_actions[player]) {
            int cnt = p.ext_counts[player] ? p.ext_counts[player][gi] : p.ext_maxk[player];
            if (cnt > p.ext_maxk[player]) cnt = p.ext_maxk[player];
            int n0 = pn;
            pn = decode_external(g, player, pn, p.ext_actions[player] + gi * p.ext_stride[player], cnt,
                                 p.ext_format[player], p.ext_fill[player], 2 * p.max_range + 1);
            if (p.safe) legality_pass(g, n0, pn);
        }
        return pn;
    }
    if (POL_IS_SCRIPTED(pol) && p.scripted) {
        int n0 = pn;
        pn = policy_scripted(g, player, pol, p.pathfinder[player], pn);
        legality_pass(g, n0, pn); // the list goes through issueSafe (Game.java:136-137), which may replace desires by NONE
    }
    return pn; // PASSIVE
}

// One decision point of RandomBiasedAI players: scan, then each deciding player's pass.  Returns the earliest completion
// time of any assignment afterwards.  simul: Game.start semantics (both lists built on the pre-issue state, issued p0 then
// p1); otherwise NaiveMCTS.simulate semantics (player 1 decides on the state that already holds player 0's actions).
DEV int rb_decide(Game &g, int n, int time, int polmask, bool simul, unsigned &decisions) {
    int par0 = 0, par1 = 0, mr, cnt0, cnt1;
    int idle = rb_scan(g, n, par0, par1, mr, cnt0, cnt1) & polmask;
    if (idle) {
        RbCtx c;
        c.time = time; c.seq_base = (uint32_t)g.hdr()[H_NEXTSEQ]; c.s0 = hdr_rng(g, H_RNGP_LO); c.k = 0; c.minr = mr; c.cancelled = false;
        #pragma unroll 1
        for (int pl = 1; pl <= 2; pl++) {
            if (!(idle & pl)) continue;
            if (!simul && pl == 2 && (idle & 1)) { int mr2, x0; rb_scan(g, n, par0, par1, mr2, x0, cnt1); } // player 0's new PRODUCEs are in flight now
            c.par0 = par0; c.par1 = par1;
            rb_player(g, pl, pl == 1 ? g.list() : g.base() + g.o_pslot, pl == 1 ? cnt0 : cnt1, simul, c);
        }
        mr = c.minr;
        if (__ballot_sync(FULLM, c.cancelled)) {
            if (!g.slim) {
                #pragma unroll 1
                for (int i = g.lane; i < g.pcw; i += 32) ((uint32_t *)g.claim())[i] = 0;
            } else { // drop the claim codes from resv[] (0xFF never occurs there: a wall is never a target)
                #pragma unroll 1
                for (int i = g.lane; i < g.pcw; i += 32) {
                    uint32_t v = ((uint32_t *)g.resv())[i], hit = 0;
                    #pragma unroll
                    for (int b = 0; b < 4; b++) if (((v >> (8 * b)) & 0xff) > (uint32_t)MRTS_MAX_CAP) hit |= 0xffu << (8 * b);
                    if (hit) ((uint32_t *)g.resv())[i] = v & ~hit;
                }
            }
        }
        if (g.lane == 0) { g.hdr()[H_NEXTSEQ] = (int32_t)(c.seq_base + c.k); hdr_set_rng(g, H_RNGP_LO, lcg_jump(g, c.s0, c.k)); }
        __syncwarp();
        decisions += c.k;
    }
    return warp_min_time(mr);
}

// The loop shared by Game.start (rts/Game.java:126-140, simul = true) and NaiveMCTS.simulate
// (ai/mcts/naivemcts/NaiveMCTS.java:297-308, simul = false) for RandomBiasedAI / PassiveAI players under CANCEL_BOTH:
// decide, then cycle.  Time jumps straight to the next completion time: while no unit is idle nothing can change
// (GameState.getNextChangeTime, GameState.java:532-546).  Returns gameover; `time` ends at the last executed cycle or tlimit.
DEV bool play_rb(Game &g, int polmask, bool simul, int tlimit, int &time, int &winner, unsigned &decisions, unsigned &ucyc) {
    bool over = game_over(g, winner); // a state that is already over ends at the very next cycle()
    int n = g.hdr()[H_NUNITS];
    #pragma unroll 1
    while (time < tlimit) {
        int mrt = rb_decide(g, n, time, polmask, simul, decisions);
        int tn = time + 1; if (!over && mrt > tn) tn = mrt;
        if (tn > tlimit) { ucyc += (unsigned)(n * (tlimit - time)); time = tlimit; break; }
        ucyc += (unsigned)(n * (tn - time));
        time = tn;
        if (cycle_execute(g, tn) > 0) over = game_over(g, winner);
        n = g.hdr()[H_NUNITS];
        if (over) return true;
    }
    return false;
}

// Game.start loop body for device policies RandomBiasedAI / PassiveAI under CANCEL_BOTH (the fast path of run_game below)
DEV void run_game_fast(Game &g, const StepParams &p, WarpStats &ws) {
    int status = g.hdr()[H_STATUS];
    if (status & ST_OVER) return;
    int t0 = g.hdr()[H_TIME], time = t0;
    int tlimit = t0 + p.n_cycles; if (tlimit > p.max_cycles) tlimit = p.max_cycles;
    int polmask = (p.policy[0] == POL_RANDOM_BIASED ? 1 : 0) | (p.policy[1] == POL_RANDOM_BIASED ? 2 : 0);
    int winner;
    unsigned decisions = 0, ucyc = 0; // per step and game: at most units x cycles, far below 2^32
    if (play_rb(g, polmask, true, tlimit, time, winner, decisions, ucyc)) {
        status |= ST_OVER | ST_COUNTED | ((winner + 1) << 8);
        stat_add(ws, g.lane, STAT_FINISHED, 1); if (winner == 0) stat_add(ws, g.lane, STAT_WINS0, 1); else if (winner == 1) stat_add(ws, g.lane, STAT_WINS1, 1); else stat_add(ws, g.lane, STAT_DRAWS, 1);
    }
    if (!(status & ST_COUNTED) && time >= p.max_cycles) { // hit the cycle cap: a draw (winner() == -1)
        status |= ST_COUNTED;
        stat_add(ws, g.lane, STAT_FINISHED, 1); stat_add(ws, g.lane, STAT_DRAWS, 1);
    }
    __syncwarp();
    if (g.lane == 0) { g.hdr()[H_TIME] = time; g.hdr()[H_STATUS] = status; }
    __syncwarp();
    stat_add(ws, g.lane, STAT_CYCLES, (unsigned long long)(time - t0));
    stat_add(ws, g.lane, STAT_DECISIONS, decisions);
    stat_add(ws, g.lane, STAT_UNIT_CYCLES, ucyc);
}

// ---- step facts for the reward functions (src/ai/reward/*.java) ----------------------------------------------------------
// The reference computes rewards from the TraceEntry of the step: the PlayerActions as issueSafe left them (illegal actions
// already replaced by NONE; conflicts NOT applied) and the PhysicalGameState before the cycle.  Counts per player:
// [0] HARVEST, [1] RETURN (ResourceGatherRewardFunction), [2] ATTACKs on a cell held by the opponent, [3] on an own unit
// (AttackRewardFunction), [4] PRODUCE Worker, [5] PRODUCE Barracks/Base, [6] PRODUCE Light/Heavy/Ranged
// (Produce*RewardFunction).  Accumulated over the decision points of the step.
DEVN void info_count(Game &g, const StepParams &p, int player, int from, int to, int32_t *o) {
    int c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0, c6 = 0;
    #pragma unroll 1
    for (int k = from + g.lane; k < to; k += 32) {
        uint32_t A0 = g.pa0()[k];
        int at = a_type(A0);
        if (at == ACT_HARVEST) c0++;
        else if (at == ACT_RETURN) c1++;
        else if (at == ACT_ATTACK) {
            int ax = (A0 >> 16) & 0xff, ay = A0 >> 24;
            if (ax < g.W && ay < g.H) {
                int gv = g.grid()[(ay + 1) * g.P + ax + 1];
                if (gv != 0 && gv != 0xFF) { int opl = u_pl(g.w0()[gv - 1]); if (opl == 2 - player) c2++; else if (opl == player + 1) c3++; }
            }
        } else if (at == ACT_PRODUCE && a_utype(A0) < MRTS_MAX_TYPES) {
            uint32_t bit = 1u << a_utype(A0);
            if (p.tm_worker & bit) c4++;
            if (p.tm_building & bit) c5++;
            if (p.tm_combat & bit) c6++;
        }
    }
    c0 = __reduce_add_sync(FULLM, c0); c1 = __reduce_add_sync(FULLM, c1); c2 = __reduce_add_sync(FULLM, c2); c3 = __reduce_add_sync(FULLM, c3);
    c4 = __reduce_add_sync(FULLM, c4); c5 = __reduce_add_sync(FULLM, c5); c6 = __reduce_add_sync(FULLM, c6);
    if (g.lane == 0) { o[0] += c0; o[1] += c1; o[2] += c2; o[3] += c3; o[4] += c4; o[5] += c5; o[6] += c6; }
    __syncwarp();
}
// CloserToEnemyBaseRewardFunction: squared distance from the opponent's first Base (list order, position BEFORE the cycle)
// to the player's closest Worker/Light/Heavy/Ranged.  before: find the base, o[7] = exists, o[11] = x | y << 8, o[8] = d2 or
// -1; after: o[9] = d2 or -1 against the remembered base position.  Also o[10] = some Resource unit still holds resources
// (ResourceGatherRewardFunction.isDone is its negation), evaluated on the state after the cycle.
DEVN void info_distance(Game &g, const StepParams &p, int player, int32_t *o, bool before) {
    int n = g.hdr()[H_NUNITS];
    int bx = 0, by = 0, exists = 0;
    if (before) {
        #pragma unroll 1
        for (int base = 0; base < n && !exists; base += 32) {
            int i = base + g.lane;
            uint32_t w = i < n ? g.w0()[i] : 0;
            bool is = i < n && u_pl(w) == 2 - player && ((p.tm_base >> u_type(w)) & 1);
            unsigned m = __ballot_sync(FULLM, is);
            if (m) { int src = __ffs(m) - 1; uint32_t bw = __shfl_sync(FULLM, w, src); bx = u_x(bw); by = u_y(bw); exists = 1; }
        }
    } else { exists = o[7]; bx = o[11] & 0xff; by = (o[11] >> 8) & 0xff; }
    int best = 0x7fffffff, resleft = 0;
    #pragma unroll 1
    for (int i = g.lane; i < n; i += 32) {
        uint32_t w = g.w0()[i];
        if (u_pl(w) == player + 1 && ((p.tm_mobile >> u_type(w)) & 1)) { int dx = bx - u_x(w), dy = by - u_y(w), d = dx * dx + dy * dy; if (d < best) best = d; }
        if (((p.tm_resource >> u_type(w)) & 1) && u_res(g.w1()[i]) > 0) resleft = 1;
    }
    best = (int)__reduce_min_sync(FULLM, (unsigned)best);
    resleft = __ballot_sync(FULLM, resleft) ? 1 : 0;
    __syncwarp();
    if (g.lane == 0) {
        if (before) { o[7] = exists; o[11] = bx | (by << 8); o[8] = (exists && best != 0x7fffffff) ? best : -1; }
        else { o[9] = (exists && best != 0x7fffffff) ? best : -1; o[10] = resleft; }
    }
    __syncwarp();
}

// Game.start loop body (rts/Game.java:126-140) with exact skipping of cycles in which nothing can happen.
DEVN void run_game(Game &g, const StepParams &p, long long gi, WarpStats &ws) {
    int status = g.hdr()[H_STATUS];
#ifdef MRTS_TU_RUSH_ONLY
    int32_t *const io = nullptr; // the lean copy carries no reward facts, no sequential issue and no general issue()
#else
    int32_t *io = p.info_out ? p.info_out + gi * (2 * MRTS_INFO_WORDS) : nullptr;
#endif
    if (io) { if (g.lane < 2 * MRTS_INFO_WORDS) io[g.lane] = 0; __syncwarp(); }
    if (status & ST_OVER) return;
    int t0 = g.hdr()[H_TIME];
    int tlimit = t0 + p.n_cycles; if (tlimit > p.max_cycles) tlimit = p.max_cycles;
    int winner;
    bool over = game_over(g, winner); // a state that is already over ends at the very next cycle()
    bool first = true;
    if (io) { info_distance(g, p, 0, io, true); info_distance(g, p, 1, io + MRTS_INFO_WORDS, true); }
    // device policies emit self-consistent lists; under CANCEL_BOTH they can be issued in parallel (issue_policy_lists)
    // (not under MRTS_FLAG_PO_POLICIES: a list built on a partial view may collide with a hidden unit's reservation)
#ifdef MRTS_TU_RUSH_ONLY
    const bool fast_issue = true, seq_issue = false;
#else
    const bool fast_issue = p.conflict == 1 && p.policy[0] != POL_EXTERNAL && p.policy[1] != POL_EXTERNAL && !p.po_policies;
    const bool seq_issue = p.sequential_issue != 0;
#endif
    unsigned long long decisions = 0, ucyc = 0;
    #pragma unroll 1
    for (;;) {
        int time = g.hdr()[H_TIME];
        if (time >= tlimit) break;
        int pn0 = run_policy(g, p, gi, 0, 0, first);
        if (io) info_count(g, p, 0, 0, pn0, io);
        if (seq_issue) issue_pending(g, 0, pn0);
        int pn1 = run_policy(g, p, gi, 1, pn0, first);
        if (io) info_count(g, p, 1, pn0, pn1, io + MRTS_INFO_WORDS);
        first = false;
        if (seq_issue) issue_pending(g, pn0, pn1);
        else if (fast_issue) issue_policy_lists(g, pn0, pn1);
        else { issue_pending(g, 0, pn0); issue_pending(g, pn0, pn1); }
        decisions += pn1;
        int mrt = min_ready_time(g);
        int tn = time + 1; if (!over && mrt > tn) tn = mrt;
#ifdef MRTS_TU_RUSH_ONLY
        if (tn > time + 1) {
#else
        if (tn > time + 1 && p.scripted) {
#endif
            // Cycles time+1 .. tn-1 are skipped because nothing can change in them, but the reference still calls getAction
            // in each.  For the scripted AIs the FIRST of those calls is not a no-op: translateActions drops the entries that
            // completed while being executed in this cycle (Train), which decides their position in the map when they are
            // re-inserted later.  Replay that one call (it cannot emit actions: no own unit is idle); the rest are no-ops.
#pragma unroll 1
            for (int pl = 0; pl < 2; pl++)
                if (POL_IS_SCRIPTED(p.policy[pl])) {
#ifndef MRTS_TU_RUSH_ONLY
                    if (p.po_policies) po_hide(g, pl);
#endif
                    policy_scripted(g, pl, p.policy[pl], p.pathfinder[pl], 0);
#ifndef MRTS_TU_RUSH_ONLY
                    if (p.po_policies) po_unhide(g);
#endif
                }
        }
        int nu = g.hdr()[H_NUNITS];
        if (tn > tlimit) { ucyc += (unsigned long long)nu * (tlimit - time); __syncwarp(); if (g.lane == 0) g.hdr()[H_TIME] = tlimit; __syncwarp(); break; }
        ucyc += (unsigned long long)nu * (tn - time);
        if (cycle_execute_ni(g, tn) > 0) over = game_over(g, winner);
        if (over) {
            if (g.lane == 0) g.hdr()[H_STATUS] = status | ST_OVER | ST_COUNTED | ((winner + 1) << 8);
            __syncwarp();
            stat_add(ws, g.lane, STAT_FINISHED, 1); if (winner == 0) stat_add(ws, g.lane, STAT_WINS0, 1); else if (winner == 1) stat_add(ws, g.lane, STAT_WINS1, 1); else stat_add(ws, g.lane, STAT_DRAWS, 1);
            break;
        }
    }
    if (io) { info_distance(g, p, 0, io, false); info_distance(g, p, 1, io + MRTS_INFO_WORDS, false); }
    int tend = g.hdr()[H_TIME];
    status = g.hdr()[H_STATUS];
    if (!(status & ST_COUNTED) && tend >= p.max_cycles) { // hit the cycle cap: a draw (winner() == -1)
        __syncwarp();
        if (g.lane == 0) g.hdr()[H_STATUS] = status | ST_COUNTED;
        __syncwarp();
        stat_add(ws, g.lane, STAT_FINISHED, 1); stat_add(ws, g.lane, STAT_DRAWS, 1);
    }
    stat_add(ws, g.lane, STAT_CYCLES, (unsigned long long)(tend - t0));
    stat_add(ws, g.lane, STAT_DECISIONS, decisions);
    stat_add(ws, g.lane, STAT_UNIT_CYCLES, ucyc);
}

// PathFinding.findPathToPositionInRange(start, targetpos, range, gs, null) as an operator (AStarPathFinding.java:52-79,
// BFSPathFinding.java:41-147, GreedyPathFinding.java:53-84): the unit standing on the query's cell, the game's current
// state, no extra resource usage.  -1 = null (no path, already in range, or no unit on the cell); range < 0 = findPath.
DEVN void pathfind_game(Game &g, const StepParams &p, long long gi) {
    const int32_t *q = p.pf_query + gi * 3;
    int cell = q[0], target = q[1], range = q[2], dir = -1;
    if (p.scripted && cell >= 0 && cell < g.W * g.H && target >= 0 && target < g.W * g.H) {
        int gv = g.grid()[(cell / g.W + 1) * g.P + cell % g.W + 1];
        if (gv != 0 && gv != 0xFF) dir = pf_find(g, p.pf_kind, gv - 1, target % g.W, target / g.W, range, 0);
    }
    __syncwarp();
    if (g.lane == 0) p.pf_out[gi] = dir;
}

// GameState.cycle() repeated until time == target (TestTracesIntegrity.java:81-85); no policies
DEVN void run_cycles_only(Game &g, int target) {
    int winner;
    bool over = game_over(g, winner);
    #pragma unroll 1
    for (;;) {
        int time = g.hdr()[H_TIME];
        if (time >= target) break;
        int mrt = min_ready_time(g);
        int tn = time + 1; if (!over && mrt > tn) tn = mrt;
        if (tn > target) { __syncwarp(); if (g.lane == 0) g.hdr()[H_TIME] = target; __syncwarp(); break; }
        if (cycle_execute_ni(g, tn) > 0) over = game_over(g, winner);
        __syncwarp();
        if (g.lane == 0) { int st = g.hdr()[H_STATUS] & ~(ST_OVER | 0x300); if (over) st |= ST_OVER | ((winner + 1) << 8); g.hdr()[H_STATUS] = st; }
        __syncwarp();
    }
}

// NaiveMCTSNode's constructor loop (ai/mcts/naivemcts/NaiveMCTSNode.java:48-53): cycle() while the game is not over and neither player
// has a unit without an assignment (GameState.canExecuteAnyAction, GameState.java:416-423)
DEVN void run_cycles_to_decision(Game &g) {
    int winner;
    #pragma unroll 1
    for (;;) {
        bool over = game_over(g, winner);
        int n = g.hdr()[H_NUNITS], idle = 0;
        #pragma unroll 1
        for (int i = g.lane; i < n; i += 32) idle |= (u_pl(g.w0()[i]) != 0 && a_type(g.a0()[i]) == (int)AT_IDLE) ? 1 : 0;
        idle = __ballot_sync(FULLM, idle) != 0;
        __syncwarp();
        if (g.lane == 0) { int st = g.hdr()[H_STATUS] & ~(ST_OVER | 0x300); if (over) st |= ST_OVER | ((winner + 1) << 8); g.hdr()[H_STATUS] = st; }
        __syncwarp();
        if (over || idle) break;
        int time = g.hdr()[H_TIME], mrt = min_ready_time(g);
        cycle_execute_ni(g, mrt > time + 1 && mrt != MRTS_NEVER ? mrt : time + 1);
    }
}

// Unit.getUnitActions (units/Unit.java:382-522) for every idle unit of out_player, as ordered lists (the order RandomBiasedAI samples
// from and PlayerActionGenerator enumerates, rts/PlayerActionGenerator.java:56-106), plus what the generator's constructor derives
// from the state: the resource usage of the assignments in flight (GameState.getResourceUsage, GameState.java:652-664).
//   hdr[8]    : choices, resources used by player 0 / 1, resources of player 0 / 1, positions used, time, bit 0 gameover | (winner + 1) << 1 |
//               bit 3 / 4: player 0 / 1 can execute an action
//   pos[]     : the linear positions x + y * W reserved by in-flight MOVE / PRODUCE assignments, unit-list order
//   choice[][4]: unit slot, Unit.ID, type | x << 8 | y << 16 | owner << 24, number of actions (the full count even when the list is cut)
//   list[][]  : action type | (direction + 1) << 4 | x << 8 | y << 16 | (unit type + 1) << 24   (x, y: attack target; NONE lasts none_duration)
DEVN void unit_actions_game(Game &g, const StepParams &p, long long gi) {
    const int n = g.hdr()[H_NUNITS], player = p.out_player, K = p.ua_max_choices, MA = p.ua_max_actions;
    int32_t *hd = p.ua_hdr + gi * 8, *pos = p.ua_pos + gi * (long long)g.cap, *ch = p.ua_choice + gi * (long long)K * 4, *ls = p.ua_list + gi * (long long)K * MA;
    int ru0, ru1, winner, nc = 0, np = 0, can = 0;
    reserved_resources(g, ru0, ru1);
    bool over = game_over(g, winner);
    const unsigned below = (1u << g.lane) - 1;
    #pragma unroll 1
    for (int base = 0; base < n; base += 32) {
        int i = base + g.lane;
        bool idle = false, uses = false, mine = false; uint32_t w = 0, A0 = 0; int A1 = 0;
        if (i < n) {
            w = g.w0()[i]; A0 = g.a0()[i]; A1 = g.a1()[i];
            idle = a_type(A0) == (int)AT_IDLE && u_pl(w) != 0;
            mine = idle && u_pl(w) == player + 1;
            uses = a_uses_cell(a_type(A0));
        }
        can |= (__ballot_sync(FULLM, idle && u_pl(w) == 1) ? 1 : 0) | (__ballot_sync(FULLM, idle && u_pl(w) == 2) ? 2 : 0);
        unsigned mu = __ballot_sync(FULLM, uses), mm = __ballot_sync(FULLM, mine);
        if (uses) { // UnitAction.resourceUsage: linear arithmetic on x + y * W (UnitAction.java:255-270)
            int q = np + __popc(mu & below), lin = u_x(w) + u_y(w) * g.W;
            if ((unsigned)A1 < 4u) lin += (A1 == 0 ? -g.W : (A1 == 1 ? 1 : (A1 == 2 ? g.W : -1)));
            if (q < g.cap) pos[q] = lin;
        }
        np += __popc(mu);
        if (mine) {
            int c = nc + __popc(mm & below);
            if (c < K) {
                Enum e; enumerate(g, i, e);
                int cnt = e.nb + e.nfree * (e.n_aff + ((e.fl & UF_MOVE) ? 1 : 0)) + 1;
                ch[c * 4] = i; ch[c * 4 + 1] = (int32_t)g.uid()[i]; ch[c * 4 + 2] = (int32_t)(u_type(w) | (u_x(w) << 8) | (u_y(w) << 16) | (u_pl(w) << 24)); ch[c * 4 + 3] = cnt;
                #pragma unroll 1
                for (int k = 0; k < cnt && k < MA; k++) {
                    uint32_t P0; int P1, tc, cost;
                    pick_action(g, e, k, p.ua_none_duration, P0, P1, tc, cost);
                    int at = a_type(P0);
                    uint32_t v = (uint32_t)at;
                    if (at >= ACT_MOVE && at <= ACT_PRODUCE) v |= (uint32_t)(P1 + 1) << 4;
                    if (at == ACT_ATTACK) v |= ((P0 >> 16) & 0xffu) << 8 | (P0 >> 24) << 16;
                    if (at == ACT_PRODUCE) v |= (a_utype(P0) + 1u) << 24;
                    ls[c * MA + k] = (int32_t)v;
                }
            }
        }
        nc += __popc(mm);
    }
    __syncwarp();
    if (g.lane == 0) {
        hd[0] = nc; hd[1] = ru0; hd[2] = ru1; hd[3] = g.hdr()[H_RES0]; hd[4] = g.hdr()[H_RES1]; hd[5] = np; hd[6] = g.hdr()[H_TIME];
        hd[7] = (over ? 1 : 0) | ((winner + 1) << 1) | (can << 3);
    }
}

// issueSafe / issue of one staged PlayerAction, no cycle
DEVN void run_issue_only(Game &g, const StepParams &p, long long gi) {
    int pl = p.issue_player;
    int cnt = p.ext_counts[pl] ? p.ext_counts[pl][gi] : p.ext_maxk[pl];
    if (cnt > p.ext_maxk[pl]) cnt = p.ext_maxk[pl];
    int pn = decode_external(g, pl, 0, p.ext_actions[pl] + gi * p.ext_stride[pl], cnt, p.ext_format[pl],
                             p.ext_fill[pl], 2 * p.max_range + 1);
    if (p.safe) legality_pass(g, 0, pn);
    issue_pending(g, 0, pn);
}

// ---- GameState.getVectorObservation(player) (GameState.java:922-968), fully observable -----------------------------------
// Planes: hp, resources, owner ((owner + player) % 2 + 1), type + 1, current action type, terrain; out = [C=6][H][W] of
// this game, u8 or i32.  Almost every cell is empty, so the planes are first written in bulk with 16-byte stores (zeros;
// the map's terrain plane for plane 5) and the units' five values are then scattered over them: the warp barrier orders
// the two writes and the second one merges in L2, so DRAM sees each output byte once.  w0/w1/a0 point at the unit table
// (shared memory inside the step kernels, HBM in k_observe).
// ---- bulk stores shared -> global (cp.async.bulk, the 1-D TMA path): one instruction by one lane moves up to 4 KB ----------------
// The fused step + observation kernel writes the planes' and masks' zeros from a block of zeros in shared memory this way: almost
// every output byte is a zero, so the LSU only sees the few values scattered over them afterwards.
#ifdef MRTS_EMU
DEV void bulk_zero(uint32_t, int, void *dst, size_t bytes) { memset(dst, 0, bytes); }
DEV void bulk_copy(uint32_t, void *, uint32_t) {}
DEV void bulk_wait_all() {}
#else
// lane 0 only: dst 16-byte aligned, bytes a multiple of 16; zsm = shared-window address of zbytes zeros
DEV void bulk_zero(uint32_t zsm, int zbytes, void *dst, size_t bytes) {
    char *d = (char *)dst;
    #pragma unroll 1
    for (size_t o = 0; o < bytes; o += (size_t)zbytes) {
        uint32_t sz = bytes - o < (size_t)zbytes ? (uint32_t)(bytes - o) : (uint32_t)zbytes;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(d + o), "r"(zsm), "r"(sz) : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
DEV void bulk_copy(uint32_t src_sm, void *dst, uint32_t bytes) { // lane 0 only; joins the bulk group the next bulk_zero commits, or commit below
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_sm), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// every bulk store of this thread has been performed: later stores to the same bytes land on top of them
DEV void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
#endif

// prezeroed: planes 0-4 were zeroed by bulk stores issued by lane 0 earlier; they are waited for before the units are scattered
DEV void obs_emit(const uint32_t *w0, const uint32_t *w1, const uint32_t *a0, int n, int W, int H, const uint8_t *terrain, int player,
                  int dtype, void *out, int lane, bool prezeroed = false, bool terrain_done = false) {
    int cells = W * H;
    if (dtype == 0) {
        uint8_t *o = (uint8_t *)out;
        if ((cells & 15) == 0) {
            int nq = cells >> 4;
            uint4 z; z.x = z.y = z.z = z.w = 0;
            if (!prezeroed) {
                #pragma unroll 4
                for (int q = lane; q < 5 * nq; q += 32) ((uint4 *)o)[q] = z;
            }
            #pragma unroll 1
            for (int q0 = lane; q0 < (terrain_done ? 0 : nq); q0 += 256) { // the terrain plane: eight 16-byte loads in flight per lane, then the stores
                uint4 t[8];
                #pragma unroll
                for (int j = 0; j < 8; j++) if (q0 + j * 32 < nq) t[j] = ((const uint4 *)terrain)[q0 + j * 32];
                #pragma unroll
                for (int j = 0; j < 8; j++) if (q0 + j * 32 < nq) ((uint4 *)(o + 5 * cells))[q0 + j * 32] = t[j];
            }
            if (prezeroed && lane == 0) bulk_wait_all();
        } else {
            #pragma unroll 1
            for (int q = lane; q < 5 * cells; q += 32) o[q] = 0;
            #pragma unroll 1
            for (int q = lane; q < cells; q += 32) o[5 * cells + q] = terrain[q];
        }
        __syncwarp();
        #pragma unroll 1
        for (int i = lane; i < n; i += 32) {
            uint32_t w = w0[i], v1 = w1[i];
            int at = a_type(a0[i]), pl = u_pl(w);
            uint8_t *c = o + u_y(w) * W + u_x(w);
            c[0] = (uint8_t)u_hp(v1); c[cells] = (uint8_t)u_res(v1);
            c[2 * cells] = (uint8_t)(pl != 0 ? ((pl - 1 + player) % 2) + 1 : 0);
            c[3 * cells] = (uint8_t)(u_type(w) + 1);
            c[4 * cells] = (uint8_t)(at == (int)AT_IDLE ? 0 : at);
        }
    } else {
        int32_t *o = (int32_t *)out;
        if ((cells & 3) == 0) {
            int nq = cells >> 2;
            int4 z; z.x = z.y = z.z = z.w = 0;
            #pragma unroll 4
            for (int q = lane; q < 5 * nq; q += 32) ((int4 *)o)[q] = z;
            #pragma unroll 2
            for (int q = lane; q < nq; q += 32) {
                uint32_t t = ((const uint32_t *)terrain)[q];
                int4 v; v.x = t & 0xff; v.y = (t >> 8) & 0xff; v.z = (t >> 16) & 0xff; v.w = t >> 24;
                ((int4 *)(o + 5 * cells))[q] = v;
            }
        } else {
            #pragma unroll 1
            for (int q = lane; q < 5 * cells; q += 32) o[q] = 0;
            #pragma unroll 1
            for (int q = lane; q < cells; q += 32) o[5 * cells + q] = terrain[q];
        }
        __syncwarp();
        #pragma unroll 1
        for (int i = lane; i < n; i += 32) {
            uint32_t w = w0[i], v1 = w1[i];
            int at = a_type(a0[i]), pl = u_pl(w);
            int32_t *c = o + u_y(w) * W + u_x(w);
            c[0] = u_hp(v1); c[cells] = u_res(v1);
            c[2 * cells] = pl != 0 ? ((pl - 1 + player) % 2) + 1 : 0;
            c[3 * cells] = u_type(w) + 1;
            c[4 * cells] = at == (int)AT_IDLE ? 0 : at;
        }
    }
    __syncwarp();
}
DEV size_t obs_bytes_per_game(int W, int H, int C, int dtype) { return (size_t)C * W * H * (dtype == 0 ? 1 : 4); }
DEV size_t mask_bytes_per_game(int W, int H, int K, int dtype) { return (size_t)W * H * (dtype == 2 ? (size_t)((K + 7) >> 3) : (size_t)K * (dtype == 0 ? 1 : 4)); }
// the map blob's terrain plane (layout.h)
DEV const uint8_t *map_terrain(const uint32_t *blob, int W, int H, int cap) { return (const uint8_t *)(blob + mrts_map_terrain_offset_words(W, H, cap)); }

// ---- GameState.getVectorObservation (GameState.java:922-968) / PartiallyObservableGameState (:35-71,82-179) ------------
// Planes: hp, resources, owner ((owner+player)%2+1), type+1, current action type, terrain [, my visibility, visible
// enemies' visibility].  out = [n_games][C][H][W], u8 or i32.  resv/claim are reused as the two visibility maps.
DEV void stamp_sight(Game &g, uint8_t *map, uint32_t w) {
    int r = ut_sight(g, u_type(w)), d = 2 * r + 1, x0 = u_x(w) - r, y0 = u_y(w) - r;
    #pragma unroll 1
    for (int q = g.lane; q < d * d; q += 32) {
        int dx = q % d - r, dy = q / d - r, x = x0 + q % d, y = y0 + q / d;
        if (x >= 0 && x < g.W && y >= 0 && y < g.H && dx * dx + dy * dy <= r * r) map[(y + 1) * g.P + x + 1] = 1;
    }
}
DEV void cell_planes(const Game &g, int cell, int player, bool po, int v[8]) {
    int x = cell % g.W, y = cell / g.W, pc = (y + 1) * g.P + x + 1;
    int gv = g.grid()[pc];
    #pragma unroll 1
    for (int k = 0; k < 8; k++) v[k] = 0;
    v[5] = g.grid_tmpl ? (int)(((const uint8_t *)g.grid_tmpl)[pc] == 0xFF) : 0;
    if (po) { v[6] = g.resv()[pc]; v[7] = g.claim()[pc]; }
    if (gv != 0 && gv != 0xFF) {
        int s = gv - 1;
        uint32_t w = g.w0()[s], w1 = g.w1()[s];
        int pl = u_pl(w);
        if (po && pl != player + 1 && !g.resv()[pc]) return; // not observable: removed from the observer's view
        v[0] = u_hp(w1); v[1] = u_res(w1);
        if (pl != 0) v[2] = ((pl - 1 + player) % 2) + 1;
        v[3] = u_type(w) + 1;
        int at = a_type(g.a0()[s]);
        v[4] = at == (int)AT_IDLE ? 0 : at;
    }
}
DEVN void observe_game(Game &g, const StepParams &p, long long gi) {
    int cells = g.W * g.H, C = p.partial_obs ? 8 : 6, player = p.out_player;
    bool po = p.partial_obs != 0;
    if (!po) {
        obs_emit(g.w0(), g.w1(), g.a0(), g.hdr()[H_NUNITS], g.W, g.H, map_terrain(g.grid_tmpl, g.W, g.H, g.cap), player, p.out_dtype,
                 (char *)p.out + (size_t)gi * obs_bytes_per_game(g.W, g.H, 6, p.out_dtype), g.lane);
        return;
    }
    if (po) {
        int n = g.hdr()[H_NUNITS];
        #pragma unroll 1
        for (int i = g.lane; i < g.pcw; i += 32) { ((uint32_t *)g.resv())[i] = 0; ((uint32_t *)g.claim())[i] = 0; }
        __syncwarp();
        #pragma unroll 1
        for (int i = 0; i < n; i++) { uint32_t w = g.w0()[i]; if (u_pl(w) == player + 1) stamp_sight(g, g.resv(), w); }
        __syncwarp();
        #pragma unroll 1
        for (int i = 0; i < n; i++) { // enemy units that survive the filter (PartiallyObservableGameState.java:44-53)
            uint32_t w = g.w0()[i];
            if (u_pl(w) != 0 && u_pl(w) != player + 1 && g.resv()[cell_of(g, w)]) stamp_sight(g, g.claim(), w);
        }
        __syncwarp();
    }
    size_t base = (size_t)gi * C * cells;
    if ((cells & 3) == 0) {
        #pragma unroll 1
        for (int q = g.lane; q < cells / 4; q += 32) {
            int v[4][8];
            #pragma unroll 1
            for (int j = 0; j < 4; j++) cell_planes(g, q * 4 + j, player, po, v[j]);
            #pragma unroll 1
            for (int k = 0; k < C; k++) {
                if (p.out_dtype == 0)
                    ((uint32_t *)((uint8_t *)p.out + base + (size_t)k * cells))[q] =
                        (uint32_t)(v[0][k] & 0xff) | ((uint32_t)(v[1][k] & 0xff) << 8) | ((uint32_t)(v[2][k] & 0xff) << 16) | ((uint32_t)(v[3][k] & 0xff) << 24);
                else {
                    int32_t *o = (int32_t *)p.out + base + (size_t)k * cells + q * 4;
                    o[0] = v[0][k]; o[1] = v[1][k]; o[2] = v[2][k]; o[3] = v[3][k];
                }
            }
        }
    } else {
        #pragma unroll 1
        for (int cell = g.lane; cell < cells; cell += 32) {
            int v[8];
            cell_planes(g, cell, player, po, v);
            #pragma unroll 1
            for (int k = 0; k < C; k++) {
                if (p.out_dtype == 0) ((uint8_t *)p.out)[base + (size_t)k * cells + cell] = (uint8_t)v[k];
                else ((int32_t *)p.out)[base + (size_t)k * cells + cell] = v[k];
            }
        }
    }
}

// ---- JNIGridnetClient.getMasks (tests/JNIGridnetClient.java:210-223) + UnitAction.getValidActionArray
// (UnitAction.java:711-751).  The output must be zero-filled by the caller; only rows of idle own units are written.
DEVN void masks_game(Game &g, const StepParams &p, long long gi) {
    int n = g.hdr()[H_NUNITS], player = p.out_player;
    int R = 2 * p.max_range + 1, ctr = R / 2, nT = p.n_types, K = 1 + 6 + 16 + nT + R * R;
    { // The game's whole mask block is zero-filled here (16-byte stores when it is aligned), then the rows of the player's idle
      // units are written over it: the warp barrier orders the two writes and the second one merges in L2, as in obs_emit.
        size_t per_game = (size_t)g.W * g.H * (p.out_dtype == 2 ? (size_t)((K + 7) >> 3) : (size_t)K * (p.out_dtype == 0 ? 1 : 4));
        char *o = (char *)p.out + (size_t)gi * (p.out_stride > 1 ? p.out_stride : 1) * per_game;
        if ((per_game & 15) == 0 && (((size_t)p.out) & 15) == 0) {
            uint4 z; z.x = z.y = z.z = z.w = 0;
            #pragma unroll 4
            for (size_t q = g.lane; q < (per_game >> 4); q += 32) ((uint4 *)o)[q] = z;
        } else {
            #pragma unroll 1
            for (size_t q = g.lane; q < per_game; q += 32) o[q] = 0;
        }
        __syncwarp();
    }
    #pragma unroll 1
    for (int s = 0; s < n; s++) {
        uint32_t w = g.w0()[s];
        if (u_pl(w) != player + 1 || a_type(g.a0()[s]) != AT_IDLE) continue; // uniform across lanes
        Enum e; enumerate(g, s, e);
        bool mv = (e.fl & UF_MOVE) != 0;
        int pr_m = e.n_aff > 0 ? e.free_m : 0, mv_m = mv ? e.free_m : 0;
        size_t cell = (size_t)gi * (p.out_stride > 1 ? p.out_stride : 1) * g.W * g.H + (size_t)u_y(w) * g.W + u_x(w), row = cell * K;
        int MB = (K + 7) >> 3; // bit-packed row length in bytes (out_dtype 2)
        #pragma unroll 1
        for (int jb = 0; jb < K; jb += 32) {
            int j = jb + g.lane;
            int v = 0;
            if (j >= K) v = 0;
            else if (j == 0) v = 1;
            else if (j < 7) {
                switch (j - 1) {
                    case ACT_NONE: v = 1; break;
                    case ACT_MOVE: v = mv_m != 0; break;
                    case ACT_HARVEST: v = e.harv_m != 0; break;
                    case ACT_RETURN: v = e.ret_m != 0; break;
                    case ACT_PRODUCE: v = pr_m != 0; break;
                    case ACT_ATTACK: v = e.n_atk > 0; break;
                }
            } else if (j < 11) v = (mv_m >> (j - 7)) & 1;
            else if (j < 15) v = (e.harv_m >> (j - 11)) & 1;
            else if (j < 19) v = (e.ret_m >> (j - 15)) & 1;
            else if (j < 23) v = (pr_m >> (j - 19)) & 1;
            else if (j < 23 + nT) {
                int ut = j - 23, np = ut_nprod(g, e.t);
                if (e.nfree > 0) for (int k = 0; k < np; k++) if (ut_prod(g, e.t, k) == ut && ((e.aff_m >> k) & 1)) v = 1;
            } else if (e.fl & UF_ATTACK) {
                int r = j - 23 - nT, ax = u_x(w) + r % R - ctr, ay = u_y(w) + r / R - ctr;
                if (ax >= 0 && ay >= 0 && ax < g.W && ay < g.H) {
                    int gv = g.grid()[(ay + 1) * g.P + ax + 1];
                    if (gv != 0 && gv != 0xFF) v = enemy_in_range(g, w, g.w0()[gv - 1], e.range * e.range) ? 1 : 0;
                }
            }
            if (p.out_dtype == 2) { // bit j of the row = element j: one ballot packs 32 elements, lanes 0..3 store its bytes
                unsigned bits = __ballot_sync(FULLM, v != 0);
                int byte = (jb >> 3) + g.lane;
                if (g.lane < 4 && byte < MB) ((uint8_t *)p.out)[cell * MB + byte] = (uint8_t)(bits >> (8 * g.lane));
            } else if (j < K) {
                if (p.out_dtype == 0) ((uint8_t *)p.out)[row + j] = (uint8_t)v;
                else ((int32_t *)p.out)[row + j] = v;
            }
        }
    }
}

// The same masks, bit-packed, written by the step kernels for the state they leave behind (mrts_batch_set_mask_outputs): one lane
// per idle unit builds the unit's whole row (K <= 128 bits) in registers from the category masks of enumerate().  The game's block
// is zeroed first (16-byte stores, or bulk stores issued earlier when `prezeroed`).
DEV void masks_emit_bits(const Game &g, const StepParams &p, int player, void *out, bool prezeroed) {
    const int n = g.hdr()[H_NUNITS], R = 2 * p.max_range + 1, ctr = R / 2, nT = p.n_types, K = 1 + 6 + 16 + nT + R * R, MB = (K + 7) >> 3;
    const size_t per_game = (size_t)g.W * g.H * MB;
    uint8_t *o = (uint8_t *)out;
    if (!prezeroed) {
        if ((per_game & 15) == 0) {
            uint4 z; z.x = z.y = z.z = z.w = 0;
            #pragma unroll 4
            for (size_t q = g.lane; q < (per_game >> 4); q += 32) ((uint4 *)o)[q] = z;
        } else {
            #pragma unroll 1
            for (size_t q = g.lane; q < per_game; q += 32) o[q] = 0;
        }
    } else if (g.lane == 0) bulk_wait_all();
    __syncwarp();
    #pragma unroll 1
    for (int s = g.lane; s < n; s += 32) {
        uint32_t w = g.w0()[s];
        if (u_pl(w) != player + 1 || a_type(g.a0()[s]) != AT_IDLE) continue;
        Enum e; enumerate(g, s, e);
        const int mv_m = (e.fl & UF_MOVE) ? e.free_m : 0, pr_m = e.n_aff > 0 ? e.free_m : 0;
        unsigned long long lo = 3ull /* the unit can act; NONE */ | (mv_m ? 4ull : 0) | (e.harv_m ? 8ull : 0) | (e.ret_m ? 16ull : 0) | (pr_m ? 32ull : 0) |
                                (e.n_atk > 0 ? 64ull : 0) | ((unsigned long long)mv_m << 7) | ((unsigned long long)e.harv_m << 11) |
                                ((unsigned long long)e.ret_m << 15) | ((unsigned long long)pr_m << 19), hi = 0;
        if (e.nfree > 0) {
            int np = ut_nprod(g, e.t);
            #pragma unroll 1
            for (int k = 0; k < np; k++) if ((e.aff_m >> k) & 1) lo |= 1ull << (23 + ut_prod(g, e.t, k));
        }
        if (e.fl & UF_ATTACK) {
            const int base = 23 + nT;
            if (e.range == 1) {
                #pragma unroll
                for (int d = 0; d < 4; d++)
                    if ((e.atk_m >> d) & 1) { int j = base + (ddy(d) + ctr) * R + ddx(d) + ctr; if (j < 64) lo |= 1ull << j; else hi |= 1ull << (j - 64); }
            } else {
                const int sq = e.range * e.range;
                #pragma unroll 1
                for (int i = 0; i < n; i++) {
                    uint32_t ow = g.w0()[i];
                    if (enemy_in_range(g, w, ow, sq)) {
                        int j = base + (u_y(ow) - u_y(w) + ctr) * R + (u_x(ow) - u_x(w)) + ctr;
                        if (j < 64) lo |= 1ull << j; else hi |= 1ull << (j - 64);
                    }
                }
            }
        }
        uint8_t *row = o + ((size_t)u_y(w) * g.W + u_x(w)) * MB;
        #pragma unroll 1
        for (int b = 0; b < MB; b++) row[b] = (uint8_t)((b < 8 ? lo >> (8 * b) : hi >> (8 * (b - 8))) & 0xff);
    }
    __syncwarp();
}

// ---- NaiveMCTS playout (ai/mcts/naivemcts/NaiveMCTS.java:195-223,297-308) -----------------------------------------------
// PartiallyObservableGameState(gs, observer) (rts/PartiallyObservableGameState.java:35-71): drop every unit that is not
// the observer's and lies outside the sight radius of all observer units (with its assignment).
DEV void po_filter(Game &g, int observer) {
    int n = g.hdr()[H_NUNITS];
#pragma unroll 1
    for (int i = g.lane; i < g.pcw; i += 32) ((uint32_t *)g.claim())[i] = 0;
    __syncwarp();
#pragma unroll 1
    for (int i = 0; i < n; i++) { uint32_t w = g.w0()[i]; if (u_pl(w) == observer + 1) stamp_sight(g, g.claim(), w); }
    __syncwarp();
    int removed = 0;
#pragma unroll 1
    for (int base = 0; base < n; base += 32) {
        int i = base + g.lane;
        bool drop = false;
        if (i < n) { uint32_t w = g.w0()[i]; drop = u_pl(w) != observer + 1 && !g.claim()[cell_of(g, w)]; }
        if (drop) g.a0()[i] |= A0_DEAD;
        removed += __popc(__ballot_sync(FULLM, drop));
    }
    __syncwarp();
#pragma unroll 1
    for (int i = g.lane; i < g.pcw; i += 32) ((uint32_t *)g.claim())[i] = 0;
    __syncwarp();
    if (removed) compact_units(g);
}

// SimpleSqrtEvaluationFunction3.base_score (ai/evaluation/SimpleSqrtEvaluationFunction3.java:32-44) and
// SimpleEvaluationFunction.base_score (SimpleEvaluationFunction.java:27-36): float accumulation in unit-list order.
DEV float base_score(const Game &g, int fn, int player) {
    float score = __fmul_rn((float)g.hdr()[H_RES0 + player], 20.0f);
    bool any = false;
    int n = g.hdr()[H_NUNITS];
#pragma unroll 1
    for (int i = 0; i < n; i++) {
        uint32_t w = g.w0()[i];
        if (u_pl(w) != player + 1) continue;
        any = true;
        uint32_t w1 = g.w1()[i];
        int t = u_type(w), cost = ut_cost(g, t), mhp = ut_hp(g, t), hp = u_hp(w1);
        score = __fadd_rn(score, __fmul_rn((float)u_res(w1), 10.0f));
        if (fn == 0) { // score += 40f * cost * Math.sqrt(hp / maxhp): int division, double product, narrowing +=
            double v = __dmul_rn((double)__fmul_rn(40.0f, (float)cost), sqrt((double)(mhp ? hp / mhp : 0)));
            score = (float)__dadd_rn((double)score, v);
        } else score = __fadd_rn(score, __fdiv_rn(__fmul_rn(40.0f, (float)(cost * hp)), (float)mhp));
    }
    if (fn == 0 && !any) return 0.0f;
    return score;
}
DEV float evaluate_state(const Game &g, int fn, int maxplayer) {
    float s1 = base_score(g, fn, maxplayer), s2 = base_score(g, fn, 1 - maxplayer);
    if (fn == 0) {
        float sum = __fadd_rn(s1, s2);
        if (sum == 0.0f) return 0.5f;
        return __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, s1), sum), 1.0f);
    }
    return __fsub_rn(s1, s2);
}

// simulate(gs2, gs2.getTime() + depth):  do { if (gs.isComplete()) gameover = gs.cycle(); else { gs.issue(policy(0));
// gs.issue(policy(1)); } } while (!gameover && time < limit) -- issue(), not issueSafe(); player 1 samples after
// player 0's actions are in flight, so no same-cycle conflict can arise and both lists are inserted in parallel.
DEV void run_rollout(Game &g, const StepParams &p, long long r, WarpStats &ws) {
    if (p.observer >= 0) po_filter(g, p.observer);
    long long seed = p.ro_seeds ? p.ro_seeds[r] : r;
    __syncwarp();
    if (g.lane == 0) {
        hdr_set_rng(g, H_RNGP_LO, ((unsigned long long)seed ^ 0x5DEECE66DULL) & MASK48);
        hdr_set_rng(g, H_RNGC_LO, ((unsigned long long)(seed ^ 0x5851F42D4C957F2DLL) ^ 0x5DEECE66DULL) & MASK48);
        hdr_set_rng(g, H_RNGD_LO, ((unsigned long long)(seed ^ 0x14057B7EF767814FLL) ^ 0x5DEECE66DULL) & MASK48);
    }
    __syncwarp();
    int t0 = g.hdr()[H_TIME], time = t0, winner;
    unsigned decisions = 0, ucyc = 0;
    if (p.depth >= 0) play_rb(g, 3, false, t0 + p.depth, time, winner, decisions, ucyc); // depth < 0: mrts_batch_evaluate, no playout
    __syncwarp();
    if (g.lane == 0) {
        g.hdr()[H_TIME] = time;
        float ev = evaluate_state(g, p.eval_fn, p.maxplayer);
        if (p.ro_eval) p.ro_eval[r] = ev;
        if (p.ro_time) p.ro_time[r] = time - t0;
    }
    stat_add(ws, g.lane, STAT_CYCLES, (unsigned long long)(time - t0));
    stat_add(ws, g.lane, STAT_DECISIONS, decisions);
    stat_add(ws, g.lane, STAT_UNIT_CYCLES, ucyc);
    stat_add(ws, g.lane, STAT_FINISHED, 1);
}

// Bodies of the step kernels for one thread.  Three kernels share the loop below; what differs is which entry points
// are compiled in, so that the hot ones stay small enough for the instruction caches and fully inlined:
//   KERNEL_FAST    MODE_GAME with RandomBiasedAI / PassiveAI players under CANCEL_BOTH
//   KERNEL_ROLLOUT MODE_ROLLOUT
//   KERNEL_FAST_OBS KERNEL_FAST + the post-step observation planes of both players (mrts_batch_set_observation_outputs)
//   KERNEL_GENERIC everything else (external actions, scripted policies, other conflict policies, cycle-only,
//                  issue-only, observations, masks)
enum { KERNEL_FAST = 0, KERNEL_ROLLOUT = 1, KERNEL_GENERIC = 2, KERNEL_FAST_OBS = 3, N_KERNELS = 4 };
// Copies of the specialised kernels compiled for one map size and unit capacity (step_kernel_body<KERNEL, W, H, CAP>): the
// reference's small standard maps with the capacity mrts_batch_create derives for their basesWorkers layouts.
// X(kernel, W, H, capacity, minimum resident CTAs per SM)
#define MRTS_FIXED_VARIANTS(X) \
    X(KERNEL_FAST, 16, 16, 128, MRTS_MIN_BLOCKS) \
    X(KERNEL_FAST, 8, 8, 64, MRTS_MIN_BLOCKS) \
    X(KERNEL_ROLLOUT, 16, 16, 128, MRTS_MIN_BLOCKS_ROLLOUT) \
    X(KERNEL_ROLLOUT, 8, 8, 64, MRTS_MIN_BLOCKS_ROLLOUT) \
    X(KERNEL_ROLLOUT, 32, 32, 252, MRTS_MIN_BLOCKS_ROLLOUT)

// the warp's next work item: the global counter hands out the items behind every warp's static first one
// (CHUNK items per draw: rollouts from a partially observable root can be a handful of cycles long, and one atomic per
// rollout on a single address would become the limit)
template <int CHUNK>
DEV long long next_item(const StepParams &p, int lane, long long first_dynamic, long long cur) {
    if (CHUNK > 1 && cur >= first_dynamic && ((cur - first_dynamic + 1) % CHUNK) != 0) return cur + 1; // still inside the drawn chunk
    unsigned long long k = 0;
    __syncwarp();
    if (lane == 0) k = atomicAdd(p.work_counter, 1ULL);
    k = __shfl_sync(FULLM, k, 0);
    return first_dynamic + (long long)k * CHUNK;
}

// FW, FH, FCAP > 0: a copy of the kernel for one fixed map size and unit capacity (a batch without scripted-policy words): the
// shared-memory layout, the padded row length and the capacity are compile-time constants there, so the offset arithmetic of
// every accessor folds into immediates
template <int KERNEL, int FW = 0, int FH = 0, int FCAP = 0>
DEV void step_kernel_body(const StepParams &p, unsigned char *smem, int tid, int nthreads, int bid, int nblocks) {
#ifdef MRTS_EMU
    mrts_smem = smem;
#endif
    uint32_t *cst = (uint32_t *)mrts_smem;
    #pragma unroll 1
    for (int i = tid; i < MRTS_CONST_WORDS; i += nthreads) cst[i] = p.cst[i];
    __syncthreads();
    constexpr bool FIXED = FW > 0;
#ifdef MRTS_TU_RUSH_ONLY
    constexpr bool LEAN = true; // MODE_GAME with scripted rushes only: no other mode, no fused outputs, no in-kernel environment reset
#else
    constexpr bool LEAN = false;
#endif
    constexpr SmemLayout LC = mrts_smem_layout(FIXED ? FW : 8, FIXED ? FH : 8, FIXED ? FCAP : 32, 0, 0, 0);
    const SmemLayout &L = FIXED ? LC : p.L;
    const int pW = FIXED ? FW : p.W, pH = FIXED ? FH : p.H, pcap = FIXED ? FCAP : p.cap, puw = FIXED ? MRTS_UNIT_WORDS_CORE : p.uw;
    int warp = tid >> 5, lane = tid & 31, wpc = nthreads >> 5;
    int region = MRTS_CONST_WORDS * 4 + warp * L.total;
#ifndef MRTS_EMU
    asm volatile("" : "+r"(lane)); // keep the lane id in a register instead of re-reading the special register
#endif
    Game g;
    g_bind(g, region, L, pW, pH, pcap, lane, p.conflict, p.scripted,
           p.scripted == 2 ? p.astar_scratch + ((long long)bid * wpc + warp) * p.astar_stride : nullptr, KERNEL == KERNEL_FAST_OBS);
    if (KERNEL == KERNEL_GENERIC && (LEAN || p.scripted)) { // pathfinding scratch: no stale marks, all buckets empty, generation 0
        #pragma unroll 1
        for (int i = lane; i < (p.W + 2) * (p.H + 2); i += 32) g.as_mark[i] = 0;
        #pragma unroll 1
        for (int i = lane; i < MRTS_ASTAR_HEADS(p.W, p.H); i += 32) g.as_head[i] = 0xFFFF;
        if (lane == 0) *g.as_gen = 0;
        __syncwarp();
    }
    // the warp's counters live in shared memory: every lane adds the same (warp-uniform) amounts to its own view of them, so
    // only lane 0's stores matter; 16 registers stay free for the game loop
    // KERNEL_FAST_OBS: the CTA's block of zeros behind the warps' regions (never written again) feeds the bulk stores
    uint32_t zsm = 0;
    if (KERNEL == KERNEL_FAST_OBS && p.zero_bytes > 0) {
        unsigned char *zb = mrts_smem + MRTS_CONST_WORDS * 4 + wpc * L.total;
        #pragma unroll 1
        for (int i = tid; i < p.zero_bytes / 16; i += nthreads) { uint4 z; z.x = z.y = z.z = z.w = 0; ((uint4 *)zb)[i] = z; }
#ifndef MRTS_EMU
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // the zeros are visible to the async proxy that reads them
#endif
        __syncthreads();
        zsm = smem_window(MRTS_CONST_WORDS * 4 + wpc * L.total);
    }
    uint32_t tsm = 0; // the terrain plane and the grid template of the batch's one map, staged once per CTA
    if (KERNEL == KERNEL_FAST_OBS && (p.terr_bytes > 0 || p.tmpl_bytes > 0)) {
        unsigned char *tb = mrts_smem + MRTS_CONST_WORDS * 4 + wpc * L.total + p.zero_bytes;
        const uint4 *ter = (const uint4 *)map_terrain(p.maps, pW, pH, pcap);
        #pragma unroll 1
        for (int i = tid; i < p.terr_bytes / 16; i += nthreads) ((uint4 *)tb)[i] = ter[i];
        #pragma unroll 1
        for (int i = tid; i < p.tmpl_bytes / 16; i += nthreads) ((uint4 *)(tb + p.terr_bytes))[i] = ((const uint4 *)p.maps)[i];
#ifndef MRTS_EMU
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
        __syncthreads();
        if (p.terr_bytes > 0) tsm = smem_window(MRTS_CONST_WORDS * 4 + wpc * L.total + p.zero_bytes);
        if (p.tmpl_bytes > 0) g.tmpl_sm = (const uint4 *)(tb + p.terr_bytes);
    }
    WarpStats &ws = *(WarpStats *)(mrts_smem + region + L.stats);
    if (lane < N_WARP_STATS) ws.v[lane] = 0;
    __syncwarp();
    long long n_items = KERNEL == KERNEL_ROLLOUT ? p.n_games * p.rollouts_per_game : p.n_games;
    // A warp's first item is static; the next ones come from a global counter, so a warp that drew cheap games (or rollouts
    // that ended early) takes more of them and the launch has no long tail of a few unlucky warps.
#pragma unroll 1
    for (long long item = (long long)bid * wpc + warp; item < n_items; item = next_item<KERNEL == KERNEL_ROLLOUT ? 8 : 1>(p, lane, (long long)nblocks * wpc, item)) {
        long long gi = KERNEL == KERNEL_ROLLOUT ? item / p.rollouts_per_game : item;
        const uint32_t *blob = p.maps + (size_t)(gi % p.n_maps) * p.map_words;
        g.grid_tmpl = blob;
        if (KERNEL == KERNEL_GENERIC && !LEAN) { g.ff_cache = p.ff_cache ? p.ff_cache + (size_t)gi * 2 * p.ff_stride : nullptr; g.ff_stride = p.ff_stride; }
        int32_t *ghdr = p.hdr + gi * MRTS_HDR_WORDS;
        uint32_t *gun = p.units + gi * (long long)puw * pcap;
        g_load(g, ghdr, gun, KERNEL != KERNEL_ROLLOUT && (LEAN || p.mode == MODE_GAME) && p.auto_reset, p.max_cycles);
        stat_add(ws, lane, STAT_IO_READ, (unsigned long long)(MRTS_HDR_WORDS * 4 + puw * 4 * g.hdr()[H_NUNITS]));
        if (KERNEL == KERNEL_ROLLOUT) { run_rollout(g, p, item, ws); stat_add(ws, lane, STAT_IO_WRITE, 8); continue; } // the batch itself is not modified
        int err0 = g.hdr()[H_ERR];
        // outputs of the state this step leaves behind: their zeros (planes 0-4 of each observation, the whole mask block) leave
        // now through bulk stores, under the game's cycle(s); the values are scattered over them once they have landed
        const int cells = pW * pH;
        const size_t obs_pg = obs_bytes_per_game(pW, pH, 6, p.obs_dtype);
        const size_t mask_pg = mask_bytes_per_game(pW, pH, 1 + 6 + 16 + p.n_types + (2 * p.max_range + 1) * (2 * p.max_range + 1), 2);
        const bool obs_bulk = KERNEL == KERNEL_FAST_OBS && zsm != 0 && p.obs_dtype == 0 && ((5 * cells) & 15) == 0 && (obs_pg & 15) == 0;
        const bool mask_bulk = KERNEL == KERNEL_FAST_OBS && zsm != 0 && (mask_pg & 15) == 0;
        if (KERNEL == KERNEL_FAST_OBS && lane == 0) {
            #pragma unroll 1
            for (int pl = 0; pl < 2; pl++) {
                if (obs_bulk && p.obs_out[pl]) {
                    bulk_zero(zsm, p.zero_bytes, (char *)p.obs_out[pl] + (size_t)gi * p.out_stride * obs_pg, (size_t)5 * cells);
                    if (tsm) bulk_copy(tsm, (char *)p.obs_out[pl] + (size_t)gi * p.out_stride * obs_pg + (size_t)5 * cells, (uint32_t)cells); // plane 5 never changes
                }
                if (mask_bulk && p.mask_out[pl]) bulk_zero(zsm, p.zero_bytes, (char *)p.mask_out[pl] + (size_t)gi * p.out_stride * mask_pg, mask_pg);
            }
        }
#ifdef MRTS_DBG_NO_GAME // (profiling experiments only: build_variants/, never the shipped library)
        if (KERNEL == KERNEL_FAST_OBS) {} else
#endif
        if (KERNEL == KERNEL_FAST || KERNEL == KERNEL_FAST_OBS) run_game_fast(g, p, ws);
        else if (LEAN || p.mode == MODE_GAME) run_game(g, p, gi, ws);
        else if (p.mode == MODE_CYCLE_ONLY) run_cycles_only(g, p.t_target ? p.t_target[gi] : g.hdr()[H_TIME] + p.n_cycles);
        else if (p.mode == MODE_ISSUE_ONLY) run_issue_only(g, p, gi);
        else if (p.mode == MODE_OBSERVE) { observe_game(g, p, gi); stat_add(ws, lane, STAT_IO_WRITE, obs_bytes_per_game(g.W, g.H, p.partial_obs ? 8 : 6, p.out_dtype)); continue; }
        else if (p.mode == MODE_PATHFIND) { pathfind_game(g, p, gi); continue; }
        else if (p.mode == MODE_UNIT_ACTIONS) { unit_actions_game(g, p, gi); continue; }
        else if (p.mode == MODE_CYCLE_DECISION) run_cycles_to_decision(g);
        else { masks_game(g, p, gi); stat_add(ws, lane, STAT_IO_WRITE, mask_bytes_per_game(g.W, g.H, 1 + 6 + 16 + p.n_types + (2 * p.max_range + 1) * (2 * p.max_range + 1), p.out_dtype)); continue; }
        if (g.hdr()[H_ERR] != err0) stat_add(ws, g.lane, STAT_ERRORS, 1);
        if ((LEAN || p.mode == MODE_GAME) && p.results_out) {
            // what mrts_batch_results reports (winner() / gameover(), PhysicalGameState.java:334-387), written here so that the
            // host can fetch it with a plain copy: a results kernel would have to wait for SM slots behind whatever persistent
            // kernel another batch has running
            int n = g.hdr()[H_NUNITS], c0 = 0, c1 = 0;
            #pragma unroll 1
            for (int i = lane; i < n; i += 32) { int pl = u_pl(g.w0()[i]); c0 += pl == 1; c1 += pl == 2; }
            c0 = __reduce_add_sync(FULLM, c0); c1 = __reduce_add_sync(FULLM, c1);
            int4 r;
            r.x = g.hdr()[H_TIME]; r.y = (c0 > 0 && c1 == 0) ? 0 : ((c1 > 0 && c0 == 0) ? 1 : -1); r.z = (c0 == 0 || c1 == 0) ? 1 : 0; r.w = g.hdr()[H_ERR];
            if (!LEAN && KERNEL == KERNEL_GENERIC && p.vec_reset) {
                // JNIGridnetVecClient.gameStep (src/tests/JNIGridnetVecClient.java:244-262,272-286): an environment whose first reward
                // function reports done, or that has run max_steps steps, is reset inside the same gameStep -- the step's results and
                // reward facts are the terminal ones, the observation (and masks) returned are the restarted game's.  The static
                // Random objects keep running; bit 1 of results[2] tells the host that the game restarted.
                int steps = g.hdr()[H_ENVSTEPS] + 1;
                bool resleft = p.info_out ? p.info_out[gi * (2 * MRTS_INFO_WORDS) + 10] != 0 : true;
                bool done0 = p.vec_reset == 1 ? r.z != 0 : (p.vec_reset == 2 ? !resleft : false);
                bool restart = done0 || steps >= p.vec_max_steps;
                __syncwarp();
                if (restart) {
                    r.z |= 2;
                    const int32_t *ih = (const int32_t *)(blob + L.pcw);
                    if (lane < MRTS_HDR_WORDS && !(lane >= H_RNGP_LO && lane <= H_RNGD_HI)) g.hdr()[lane] = lane == H_SPARE ? g.hdr()[lane] + 1 : ih[lane];
                    __syncwarp();
                    load_unit_words(g, blob + L.pcw + MRTS_HDR_WORDS, g.hdr()[H_NUNITS]);
                    cp_async_wait_all();
                    g_rebuild(g, true);
                } else if (lane == 0) g.hdr()[H_ENVSTEPS] = steps;
                __syncwarp();
            }
            if (lane == 0) ((int4 *)p.results_out)[gi] = r;
        }
        g_store(g, ghdr, gun);
        stat_add(ws, lane, STAT_IO_WRITE, (unsigned long long)(MRTS_HDR_WORDS * 4 + puw * 4 * g.hdr()[H_NUNITS] + (p.mode == MODE_GAME && p.results_out ? 16 : 0)));
#ifdef MRTS_DBG_NO_EMIT
        if (KERNEL == KERNEL_FAST_OBS) { if (lane == 0) bulk_wait_all(); } else
#endif
        if (KERNEL == KERNEL_FAST_OBS || (!LEAN && KERNEL == KERNEL_GENERIC && p.mode == MODE_GAME)) {
            #pragma unroll 1
            for (int pl = 0; pl < 2; pl++)
                if (p.obs_out[pl]) {
                    stat_add(ws, lane, STAT_IO_WRITE, obs_pg);
                    obs_emit(g.w0(), g.w1(), g.a0(), g.hdr()[H_NUNITS], g.W, g.H, map_terrain(blob, g.W, g.H, g.cap), pl, p.obs_dtype,
                             (char *)p.obs_out[pl] + (size_t)gi * p.out_stride * obs_pg, lane, obs_bulk, obs_bulk && tsm != 0);
                }
            #pragma unroll 1
            for (int pl = 0; pl < 2; pl++)
                if (p.mask_out[pl]) {
                    stat_add(ws, lane, STAT_IO_WRITE, mask_pg);
                    masks_emit_bits(g, p, pl, (char *)p.mask_out[pl] + (size_t)gi * p.out_stride * mask_pg, mask_bulk);
                }
        }
    }
    __syncwarp();
    if (lane < N_WARP_STATS && p.stats && ws.v[lane]) atomicAdd(&p.stats[lane < 8 ? lane : lane - 8 + MRTS_STATS_IO_SLOT], ws.v[lane]);
    if (lane == 0) { // no warp draws an item after it got here, so the last one to arrive can reset the counters
        unsigned long long left = atomicAdd(p.work_counter + 1, 1ULL);
        if (left + 1 == (unsigned long long)nblocks * wpc) { p.work_counter[0] = 0; p.work_counter[1] = 0; }
    }
}
