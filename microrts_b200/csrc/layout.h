// layout.h -- data layout shared by the device engine and the host C-ABI.
//
// One game = a 20-word header (MRTS_HDR_WORDS) + MRTS_UNIT_WORDS struct-of-arrays over `cap` unit slots, both in HBM and (while a
// warp owns the game) in shared memory.  Unit slot order IS the reference's PhysicalGameState.units list order
// (src/rts/PhysicalGameState.java:54); assignment (LinkedHashMap) order is carried by a per-unit sequence number.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define MRTS_HD __host__ __device__ inline
#define MRTS_HDC __host__ __device__ constexpr inline
#else
#define MRTS_HD static inline
#define MRTS_HDC static constexpr inline
#endif

#define MRTS_HDR_WORDS 20
enum {
    H_TIME = 0,     // GameState.time
    H_RES0 = 1,     // Player 0 resources
    H_RES1 = 2,
    H_NUNITS = 3,   // number of live unit slots
    H_NEXTSEQ = 4,  // next assignment sequence number (insertion order of GameState.unitActions)
    H_CANCELCTR = 5,// GameState.unitCancelationCounter
    H_STATUS = 6,   // bit0: cycle() returned gameover; bits 8-9: winner+1
    H_NEXTID = 7,   // next unit ID (Unit.next_ID, per game)
    H_RNGP_LO = 8, H_RNGP_HI = 9,   // util.Sampler.generator   (policy draws)
    H_RNGC_LO = 10, H_RNGC_HI = 11, // GameState.r               (CANCEL_RANDOM)
    H_RNGD_LO = 12, H_RNGD_HI = 13, // UnitAction.r              (random damage)
    H_ERR = 14,     // sticky MRTS_GE_* bits
    H_SPARE = 15,   // episode counter (auto-reset)
    H_ASEQ0 = 16,   // next insertion sequence of player 0's AbstractionLayerAI.actions map
    H_ASEQ1 = 17,
    H_ENVSTEPS = 18, // steps since the environment's last reset (JNIGridnetVecClient.envSteps; in-kernel auto-reset only)
    H_AIFLAGS = 19  // bit p: player p's CRush_V1.buildingRacks field (cRush/CRush_V1.java:64)
};

// per-unit words
enum { UW_W0 = 0, UW_W1, UW_A0, UW_A1, UW_TIS, UW_SEQ, UW_ID, UW_X0, UW_X1, MRTS_UNIT_WORDS };
#define MRTS_UNIT_WORDS_CORE 7 // words every kernel keeps in shared memory; X0/X1 only when scripted policies are enabled
// W0: type | (player+1)<<8 | x<<16 | y<<24            W1: (uint16)hp | (uint16)res<<16
// A0: atype(4) | flags(4) | utype<<8 | ax<<16 | ay<<24    A1: parameter (direction or NONE duration)
// TIS: issue time (UnitActionAssignment.time)          SEQ: assignment sequence      ID: Unit.ID (low 32 bits)
// X0/X1: the unit's entry in its owner's AbstractionLayerAI.actions map (scripted policies):
//   X0: kind(3: 0 none,1 train,2 build,3 harvest,4 attack) | completed<<3 | type<<4 | bx<<8 | by<<16 | bxNeg<<24 | aseqHi<<25
//   X1: target slot+1 (0 null, 0xFF dead object) | base slot+1 <<8 | aseqLo<<16
#define AT_IDLE 15u
#define A0_DEAD 0x10u

#define MRTS_MAX_TYPES 8
#define MRTS_UTT_WORDS 8 // per type
// U0: cost | hp<<8 | minDamage<<16 | maxDamage<<24
// U1: attackRange | sightRadius<<8 | harvestAmount<<16 | flags<<24
// U2: produceTime | moveTime<<16      U3: attackTime | harvestTime<<16     U4: returnTime | nProduces<<16
// U5: produces[0..3] bytes            U6: produces[4..7] bytes             U7: cost of produces[0..3], one byte each
enum { UF_RESOURCE = 1, UF_STOCKPILE = 2, UF_HARVEST = 4, UF_MOVE = 8, UF_ATTACK = 16 };

#define MRTS_JUMP_ENTRIES 65 // LCG skip-ahead: entry d advances 2*d steps (d nextDouble() draws)
#define MRTS_ETA_OFFSET (MRTS_MAX_TYPES * MRTS_UTT_WORDS + MRTS_JUMP_ENTRIES * 4) // word offset of the ETA table
// ETA table: u16 eta[type][8], indexed by action type (1 MOVE, 2 HARVEST, 3 RETURN = moveTime, 4 PRODUCE = produceTime OF
// THE ROW'S TYPE (look it up with the produced type), 5 ATTACK) -- UnitAction.ETA, UnitAction.java:307-329
#define MRTS_CONST_WORDS (MRTS_ETA_OFFSET + MRTS_MAX_TYPES * 4) // utt words + jump table (u64 pairs) + ETA table

#define MRTS_INFO_WORDS 12 // per player and game: step facts for the reward functions (engine.cuh: info_count / info_distance)
#define MRTS_MAX_CAP 252 // slot ids are bytes (0 empty, 0xFF wall); a multiple of 4 so that every unit word array starts 16-byte aligned (vector loads / stores)
#define MRTS_WARPS_PER_CTA 4

// engine error bits == MRTS_GE_* of the public header
enum { GE_UNIT_OVERFLOW = 1, GE_INCONSISTENT_OLDER = 2, GE_FAILED_PRODUCE = 4, GE_CELL_OCCUPIED = 8, GE_BAD_ACTION = 16 };

struct SmemLayout {
    int hdr, units, pa0, pa1, pslot, grid, kind, resv, claim, list, stats, astar, total; // byte offsets inside one game's region
    int povis, pohid;                                                            // MRTS_FLAG_PO_POLICIES: sight map of the deciding player, unit words of the units hidden from it
    int pcw;                                                                     // padded-grid size in 32-bit words
    int uws;                                                                     // unit words resident in shared memory
    int P;                                                                       // padded row length W + 2
    int uoff[MRTS_UNIT_WORDS + 1];                                               // byte offset of each unit word array (RDY is word uws - 1)
    int rdy;                                                                     // = uoff[uws - 1]
};

// Shared-memory-only per-unit word (index uw, after the words mirrored in HBM): RDY = completion time of the unit's
// in-flight assignment (issue time + ETA), MRTS_NEVER when idle.  Rebuilt whenever a game is loaded.
#define MRTS_NEVER 0x7fffffff

// cell "kind" byte (kind map): 0 empty, 0xFF wall / out of bounds, else owner+1 (bits 0-1: 0 neutral) | isResource<<2 |
// isStockpile<<3 | 0x10 -- what Unit.getUnitActions needs to know about a neighbouring cell, without touching the unit.
#define CK_UNIT 0x10

// A* / BFS / flood-fill scratch of one warp (scripted policies only), all u16; PC = padded cells (W+2)*(H+2): mark[PC] (query
// generation << 5 | direction the node was reached by << 3 | flags), next[PC] (bucket chains / FIFO queue), head[W*H + W + H + 2]
// (one LIFO bucket per f value), generation counter.  A node's parent is its cell minus the direction's offset and its
// coordinates follow from its index, so neither is stored.  Owned by the warp, not the game: it is initialised once per launch.
#define MRTS_ASTAR_HEADS(W, H) ((W) * (H) + (W) + (H) + 2)
#define MRTS_ASTAR_BYTES(W, H) ((4 * ((W) + 2) * ((H) + 2) + 2 * MRTS_ASTAR_HEADS(W, H) + 4 + 15) & ~15)
// pending = 0: the layout of the specialised kernels (fast game loop, rollouts), which fuse policy and issue and never
// stage a pending action list -- 8 bytes per unit slot less, which is what lets one more CTA fit per SM on small maps
// slim = 1: the layout of the fused step + observation kernel, which runs on large maps where the cell maps bound the occupancy: no
// kind map (a neighbour's kind is looked up through grid[] and the unit table) and no claim map (the few cells a cancelled pair
// leaves claimed are coded into resv[]) -- half the map bytes, 12 instead of 8 games in flight per SM on a 64x64 map
MRTS_HDC SmemLayout mrts_smem_layout(int W, int H, int cap, int scripted, int po_policies = 0, int pending = 1, int slim = 0) {
    SmemLayout L{};
    int pc = (W + 2) * (H + 2);
    int pcb = (pc + 15) & ~15;
    int capb = (cap + 15) & ~15;
    int o = 0;
    L.uws = (scripted ? MRTS_UNIT_WORDS : MRTS_UNIT_WORDS_CORE) + 1; // X0/X1 are resident only for scripted batches; +1: RDY
    L.P = W + 2;
    L.hdr = o; o += MRTS_HDR_WORDS * 4;
    for (int k = 0; k <= MRTS_UNIT_WORDS; k++) L.uoff[k] = o + k * cap * 4; // host-computed so the kernels see plain constants
    L.rdy = L.uoff[L.uws - 1];
    L.units = o; o += (L.uws * cap * 4 + 15) & ~15; // every section starts 16-byte aligned (vector fills of the cell maps)
    L.pa0 = o; o += pending ? ((cap * 4 + 15) & ~15) : 0;
    L.pa1 = o; o += pending ? ((cap * 4 + 15) & ~15) : 0;
    L.pslot = o; o += capb;
    L.grid = o; o += pcb;
    L.kind = o; o += slim ? 0 : pcb;
    L.resv = o; o += pcb;
    L.claim = o; o += slim ? 0 : pcb;
    L.list = o; o += capb;
    L.stats = o; o += 80; // the warp's 10 running counters (kept out of registers)
    L.povis = o; o += po_policies ? pcb : 0;
    L.pohid = o; o += po_policies ? ((cap * 4 + 15) & ~15) : 0;
    L.astar = o; o += scripted == 1 ? MRTS_ASTAR_BYTES(W, H) : 0; // scripted == 2: scratch in global memory
    L.total = (o + 15) & ~15;
    L.pcw = pcb / 4;
    return L;
}

// per-map blob in HBM (32-bit words): [grid template pcw][init header][init units 9*cap][terrain plane, W*H bytes of 0/1,
// 16-byte aligned: plane 5 of GameState.getVectorObservation]
MRTS_HD int mrts_map_terrain_offset_words(int W, int H, int cap) {
    int pcb = (((W + 2) * (H + 2)) + 15) & ~15;
    return (pcb / 4 + MRTS_HDR_WORDS + MRTS_UNIT_WORDS * cap + 3) & ~3;
}
MRTS_HD int mrts_map_blob_words(int W, int H, int cap) {
    return mrts_map_terrain_offset_words(W, H, cap) + ((W * H + 15) & ~15) / 4;
}
