// scripted.cuh -- device versions of the reference's scripted rush AIs and their pathfinders (included by engine.cuh).
//
//   AStarPathFinding   src/ai/abstraction/pathfinding/AStarPathFinding.java:52-79,104-138,175-295
//   BFSPathFinding     src/ai/abstraction/pathfinding/BFSPathFinding.java:41-147
//   AbstractionLayerAI src/ai/abstraction/AbstractionLayerAI.java:58-113 (translateActions), :143-245 (building position)
//   WorkerRush         src/ai/abstraction/WorkerRush.java:63-204        LightRush  src/ai/abstraction/LightRush.java:77-258
//   HeavyRush / RangedRush  src/ai/abstraction/HeavyRush.java, RangedRush.java: LightRush with Heavy / Ranged as the trained type
//   WorkerDefense      src/ai/abstraction/WorkerDefense.java:74-209     LightDefense  src/ai/abstraction/LightDefense.java:78-247
//   HeavyDefense / RangedDefense  the same class as LightDefense with the trained type swapped
//   GreedyPathFinding  src/ai/abstraction/pathfinding/GreedyPathFinding.java:53-84
//   POWorkerRush / POLightRush / POHeavyRush / PORangedRush  src/ai/abstraction/partialobservability/*.java:42-78 (exploration) + Move.java
//   Attack/Harvest/Build/Train.execute   src/ai/abstraction/{Attack.java:51,Harvest.java:72,Build.java:54,Train.java:48-128}
//   FloodFillPathFinding  src/ai/abstraction/pathfinding/FloodFillPathFinding.java:47-213
//   CRush_V1 / CRush_V2   src/ai/abstraction/cRush/CRush_V1.java:68-421, CRush_V2.java:69-479, RangedAttack.java:58-87, CRanged_Tactic.java:77-389
//   EMRDeterministico     src/ai/abstraction/EMRDeterministico.java:74-358
//
// The AI's per-unit abstract action (the value of AbstractionLayerAI.actions for that unit) lives in the unit's X0/X1
// words; the map's insertion order is the `aseq` field.  Unit references held by abstract actions (attack target,
// harvest target/base) are slot numbers that compact_units() re-maps; 0xFF stands for "an object no longer in the unit
// list".  The policy is a chain of order-dependent decisions (each desire changes what the next A* may step on), so its
// control flow is sequential -- but it is executed by the whole warp in lockstep on warp-uniform values (every lane reads
// the same shared-memory words), which lets the inner loops use the lanes: closest-unit searches and counts are one
// strided pass + a warp reduction, and the pathfinders evaluate the four neighbours of a node on four lanes.  State is
// written by lane 0 (or by the lane that owns the item) followed by __syncwarp().
#pragma once

enum { AA_NONE = 0, AA_TRAIN = 1, AA_BUILD = 2, AA_HARVEST = 3, AA_ATTACK = 4, AA_MOVE = 5 /* Move.java: walk to (bx, by) */,
       AA_RANGED_ATTACK = 6 /* cRush/RangedAttack.java: target + the closest own barracks in the base field */,
       AA_TACTIC = 7 /* cRush/CRanged_Tactic.java: target + the POSITIONS of `home` and `enemyBase` (bases never move, and the
                        reference keeps using a destroyed base's last position); x == 0xFF stands for null */ };
// type ids by name, as the reference looks them up (utt.getUnitType("Worker") ...): standard tables use these ids
#define UT_BASE 1
#define UT_BARRACKS 2
#define UT_WORKER 3
#define UT_LIGHT 4
#define UT_HEAVY 5
#define UT_RANGED 6
#define REF_NULL 0u
#define REF_DEAD 0xFFu

DEV int aa_kind(uint32_t X0) { return X0 & 7; }
DEV int aa_type(uint32_t X0) { return (X0 >> 4) & 0xF; }
DEV int aa_bx(uint32_t X0) { return (X0 & (1u << 24)) ? -1 : (int)((X0 >> 8) & 0xff); }
DEV int aa_by(uint32_t X0) { return (X0 >> 16) & 0xff; }
DEV uint32_t aa_seq(uint32_t X0, uint32_t X1) { return (X1 >> 16) | ((X0 >> 27) << 16); } // 21 bits
// AA_TACTIC: X0 = kind | hy[4:0] << 3 | ex << 8 | ey << 16 | hy[7:5] << 24 | seq, X1 = target | hx << 8 | seq (hx is NOT a unit
// reference: compact_units leaves the base field of this kind alone)
DEV int aa_thx(uint32_t X1) { return (X1 >> 8) & 0xff; }
DEV int aa_thy(uint32_t X0) { return (int)(((X0 >> 3) & 0x1f) | (((X0 >> 24) & 7) << 5)); }
DEV int aa_tex(uint32_t X0) { return (X0 >> 8) & 0xff; }
DEV int aa_tey(uint32_t X0) { return (X0 >> 16) & 0xff; }
DEV int aa_target(uint32_t X1) { return X1 & 0xff; }
DEV int aa_base(uint32_t X1) { return (X1 >> 8) & 0xff; }

// actions.put(u, aa): an existing key keeps its position in the LinkedHashMap.  Called by the whole warp; lane 0 writes.
DEV void aa_put(Game &g, int s, int player, int kind, int type, int bx, int by, int target, int base) {
    __syncwarp();
    if (g.lane == 0) {
        uint32_t oX0 = g.x0()[s], oX1 = g.x1()[s];
        uint32_t seq;
        if (aa_kind(oX0) != AA_NONE) seq = aa_seq(oX0, oX1);
        else seq = (uint32_t)g.hdr()[H_ASEQ0 + player]++;
        uint32_t X0 = (uint32_t)kind | ((uint32_t)type << 4) | ((uint32_t)(bx & 0xff) << 8) | ((uint32_t)(by & 0xff) << 16) | (bx < 0 ? (1u << 24) : 0u) |
                      (((seq >> 16) & 0x1fu) << 27);
        if (kind == AA_TACTIC) // type = home y, base = home x, (bx, by) = enemy base
            X0 = (uint32_t)kind | ((uint32_t)(type & 0x1f) << 3) | ((uint32_t)(bx & 0xff) << 8) | ((uint32_t)(by & 0xff) << 16) | ((uint32_t)((type >> 5) & 7) << 24) | (((seq >> 16) & 0x1fu) << 27);
        uint32_t X1 = (uint32_t)target | ((uint32_t)base << 8) | ((seq & 0xffffu) << 16);
        g.x0()[s] = X0; g.x1()[s] = X1;
    }
    __syncwarp();
}

// ---- warp-wide passes over the unit table (lane l looks at units l, l + 32, ...) ---------------------------------------------
// first unit in list order with the smallest key: key_of(i, w0[i]) returns the distance-like key, or -1 to skip the unit
template <class F> DEV int w_argmin(const Game &g, int n, F key_of, int *min_key = nullptr) {
    unsigned best = 0xFFFFFFFFu;
#pragma unroll 1
    for (int i = g.lane; i < n; i += 32) { int k = key_of(i, g.w0()[i]); if (k >= 0) { unsigned q = ((unsigned)k << 8) | (unsigned)i; if (q < best) best = q; } }
    best = __reduce_min_sync(FULLM, best);
    if (min_key) *min_key = best == 0xFFFFFFFFu ? 0 : (int)(best >> 8);
    return best == 0xFFFFFFFFu ? -1 : (int)(best & 0xff);
}
template <class F> DEV int w_count(const Game &g, int n, F pred) {
    int c = 0;
#pragma unroll 1
    for (int i = g.lane; i < n; i += 32) c += pred(i, g.w0()[i]) ? 1 : 0;
    return __reduce_add_sync(FULLM, c);
}
// visit the units satisfying pred in list order: the lanes test 32 units at a time, the warp then walks the set bits (body
// is warp-uniform and must not change what pred reads)
template <class P, class F> DEV void w_for_each(const Game &g, int n, P pred, F body) {
#pragma unroll 1
    for (int base = 0; base < n; base += 32) {
        int i = base + g.lane;
        unsigned mask = __ballot_sync(FULLM, i < n && pred(i, g.w0()[i]));
#pragma unroll 1
        while (mask) { int b = __ffs(mask) - 1; mask &= mask - 1; body(base + b); }
    }
}
template <class F> DEV int w_next(const Game &g, int n, int from, F pred) { // first unit after `from` satisfying pred, or -1
    unsigned best = 0xFFFFFFFFu;
#pragma unroll 1
    for (int i = g.lane; i < n; i += 32) if (i > from && (unsigned)i < best && pred(i, g.w0()[i])) best = (unsigned)i;
    best = __reduce_min_sync(FULLM, best);
    return best == 0xFFFFFFFFu ? -1 : (int)best;
}

// ---- pathfinding ---------------------------------------------------------------------------------------------------------
// The reference keeps `open` as an array sorted by descending f = heuristic + cost with a new node inserted behind every
// node of equal f, and pops from the end (AStarPathFinding.java:104-138,175-295): the node with the smallest f comes out
// first and, among equal f, the NEWEST one.  That is exactly one LIFO stack per f value, so the open list here is a bucket
// queue (head[f] -> chain through next[]): O(1) push and pop, same expansion order.  A node enters `open` at most once
// (inOpenOrClosed) and so is popped once; what the caller wants of the path -- its first step -- is stored when the node is
// pushed (A*: the first step of the path that reached it, handed down from its parent; BFS: the direction it was reached by,
// walked back at the end).  A neighbour's f is the expanded node's (a step towards the target) or two more (away from it), the
// heuristic being consistent: every f has the parity of the start's and never falls below it.  Per-cell state carries the query's
// generation number instead of being cleared for every query.
enum { PFF_INOC = 1, PFF_BLOCKED = 2, PFF_CLOSED = 4 };
#define PF_NONE 0xFFFFu
#define PF_GEN_LIMIT 2047 // generation numbers have 11 bits of the mark word
// Positions are indices into the wall-padded grid: out of bounds looks like a wall, so there are no bounds checks.  A node's
// coordinates follow from its index (a division by the padded row length: a constant in the fixed-layout kernels); neither its
// cost nor its parent is stored.
// the scratch arrays of one query; built from the shared-window address when the scratch is in shared memory, so that the
// accesses compile to LDS/STS instead of generic loads (pf_find<true>), from the global pointers otherwise
struct PfArr { uint16_t *mark, *next, *head, *gen; const uint8_t *grid, *resv; int P; };
template <bool SM> DEV PfArr pf_arrays(const Game &g) {
    PfArr a;
    if (SM) {
        int pc = g.P * (g.H + 2);
        a.mark = (uint16_t *)smem_ptr(g.as_sm); a.next = a.mark + pc; a.head = a.next + pc;
        a.gen = a.head + MRTS_ASTAR_HEADS(g.W, g.H);
    } else { a.mark = g.as_mark; a.next = g.as_next; a.head = g.as_head; a.gen = g.as_gen; }
    a.grid = g.grid(); a.resv = g.resv(); a.P = g.P;
    return a;
}
DEV int pf_flags(const PfArr &g, int pos, int gen) { int m = g.mark[pos]; return (m >> 5) == gen ? (m & 7) : 0; }
DEV void pf_set(const PfArr &g, int pos, int gen, int flags, int dir = 0) { g.mark[pos] = (uint16_t)((gen << 5) | (dir << 3) | flags); }
DEV int pf_off(const PfArr &g, int d) { return (d & 1) ? 2 - d : (d - 1) * g.P; } // up -P, right +1, down +P, left -1
// GameState.free (GameState.java:191-207) unless the cell is used by a desire already chosen this cycle (ru)
DEV bool pf_free(const PfArr &g, int pc, int fl) { return !(fl & PFF_BLOCKED) && g.grid[pc] == 0 && g.resv[pc] == 0; }
// the first step of the path that ends in `pos`: walk the parent links back to the start
DEV int pf_first_step(const PfArr &g, int pos, int start) {
    int d = -1;
#pragma unroll 1
    while (pos != start) { d = (g.mark[pos] >> 3) & 3; pos -= pf_off(g, d); }
    return d;
}
DEV int iabs(int v) { return v < 0 ? -v : v; }

// findPathToPositionInRange: direction of the first step of a shortest path from unit slot s to within `range` of
// (tx, ty), or -1 (null).  ru = target cells of the desires [0, nd) in the pending list.  Called by the whole warp with
// uniform arguments; the search state is uniform, the four neighbours of the expanded node are examined by lanes 0..3 and
// pushed in the reference's order (up, right, down, left): lanes that push into the same bucket chain themselves in that order
// (a later direction lands on top), all at once.
template <bool SM, int KIND = -1> // KIND >= 0: the search kind as a compile-time constant (the other one is not compiled in)
DEVN int pf_find_t(Game &g, int kind, int s, int tx, int ty, int range, int nd) {
    if (KIND >= 0) kind = KIND;
    const PfArr A = pf_arrays<SM>(g);
    const int lane = g.lane;
    __syncwarp();
    int gen = *A.gen + 1;
    __syncwarp();
    if (gen >= PF_GEN_LIMIT) { // generation numbers wrapped: forget every mark
        int pcells = A.P * (g.H + 2);
#pragma unroll 1
        for (int i = lane; i < pcells; i += 32) A.mark[i] = 0;
        gen = 1;
    }
    if (lane == 0) *A.gen = (uint16_t)gen;
    __syncwarp();
#pragma unroll 1
    for (int k = lane; k < nd; k += 32) { // desires target distinct cells, so the lanes never write the same mark
        uint32_t A0 = g.pa0()[k];
        if (a_uses_cell(a_type(A0))) {
            int pc = linear_target_cell(g, g.w0()[g.pslot()[k]], g.pa1()[k]);
            if (pc >= 0) pf_set(A, pc, gen, PFF_BLOCKED);
        }
    }
    __syncwarp();
    int sq = range * range;
    uint32_t sw = g.w0()[s];
    int sx = u_x(sw), sy = u_y(sw), start = cell_of(g, sw);
    int result = -1;
    const int dl = lane & 3, doffl = (dl & 1) ? 2 - dl : (dl - 1) * A.P, dxl = ddx(dl), dyl = ddy(dl); // this lane's direction
    const unsigned below = (1u << lane) - 1;
    if (kind == 0) { // A*
        const int f0 = iabs(sx - tx) + iabs(sy - ty); // f never falls below the start's
        int fhi = f0, fcur = f0;
        if (lane == 0) {
            pf_set(A, start, gen, pf_flags(A, start, gen) | PFF_INOC);
            A.next[start] = PF_NONE; A.head[f0] = (uint16_t)start;
        }
        __syncwarp();
#pragma unroll 1
        for (;;) {
            // smallest non-empty bucket: 32 buckets per probe
            int pos = A.head[fcur]; // the search mostly runs along one f value: look at the current bucket first
            if (pos == PF_NONE) {
#pragma unroll 1
                while (fcur <= fhi) { // every f has the parity of the start's: 32 candidate buckets per probe, two apart
                    unsigned m = __ballot_sync(FULLM, fcur + 2 * lane <= fhi && A.head[fcur + 2 * lane] != PF_NONE);
                    if (m) { fcur += 2 * (__ffs(m) - 1); break; }
                    fcur += 64;
                }
                if (fcur > fhi) break;
                pos = A.head[fcur];
            }
            int nxt = A.next[pos], fs = (A.mark[pos] >> 3) & 3; // fs: the first step of the path that reached pos (travels with the marks)
            __syncwarp(); // every lane has read the bucket head before it is popped
            if (lane == 0) A.head[fcur] = (uint16_t)nxt; // (a cell enters `open` once, so it is popped once: no closed test needed)
            int y = pos / A.P, hx = pos - y * A.P - 1 - tx, hy = y - 1 - ty; // offset from the target
            if (hx * hx + hy * hy <= sq) { result = pos == start ? -1 : fs; break; }
            // lanes 0..3: one neighbour each (addToOpen :104-138).  A step away from the target (or sideways past it) costs two more in f,
            // a step towards it leaves f as it is: f = heuristic + cost never needs the two terms themselves
            int np = pos + doffl, f = fcur + ((dxl * hx + dyl * hy >= 0) ? 2 : 0);
            bool ok = false;
            if (lane < 4) {
                int nfl = pf_flags(A, np, gen);
                ok = !(nfl & PFF_INOC) && pf_free(A, np, nfl);
                if (ok) pf_set(A, np, gen, nfl | PFF_INOC, pos == start ? dl : fs);
            }
            // the lanes that push into the same bucket are chained in direction order.  The Manhattan heuristic is consistent: a neighbour's f
            // is this node's (one step closer) or two more (one step farther), so there are at most two buckets, both known
            unsigned okm = __ballot_sync(FULLM, ok);
            if (okm) {
                unsigned m0 = __ballot_sync(FULLM, ok && f == fcur), m2 = okm & ~m0;
                unsigned same = f == fcur ? m0 : m2;
                int prev = (same & below) ? 31 - __clz(same & below) : -1;
                int prev_np = __shfl_sync(FULLM, np, prev < 0 ? 0 : prev);
                if (ok) A.next[np] = prev >= 0 ? (uint16_t)prev_np : A.head[f];
                __syncwarp(); // every chain start has read its bucket head before the heads move
                if (ok && !(same & ~below & ~(1u << lane))) A.head[f] = (uint16_t)np; // the last lane of a bucket's chain is its new head
                if (m2 && fcur + 2 > fhi) fhi = fcur + 2;
            }
            __syncwarp();
        }
        __syncwarp();
#pragma unroll 1
        for (int f = f0 + lane; f <= fhi; f += 32) A.head[f] = PF_NONE; // leave every bucket empty for the next query
        __syncwarp();
        return result;
    }
    // BFS: FIFO queue of positions in next[]; a cell is enqueued at most once (so it never wraps) and remembers the direction it
    // was reached by
    uint16_t *qpos = A.next;
    int oi = 1, orm = 0;
    if (lane == 0) { qpos[0] = (uint16_t)start; pf_set(A, start, gen, pf_flags(A, start, gen) | PFF_INOC); }
    __syncwarp();
#pragma unroll 1
    while (oi != orm) {
        int pos = qpos[orm];
        orm++;
        int mk = A.mark[pos], fl = (mk >> 5) == gen ? (mk & 7) : 0;
        __syncwarp();
        if (fl & PFF_CLOSED) continue;
        if (lane == 0) A.mark[pos] = (uint16_t)(mk | PFF_CLOSED);
        int y = pos / A.P, x = pos - y * A.P - 1; y -= 1;
        if ((x - tx) * (x - tx) + (y - ty) * (y - ty) <= sq) { __syncwarp(); result = pf_first_step(A, pos, start); break; }
        int np = pos + doffl;
        bool ok = false;
        if (lane < 4) {
            int nfl = pf_flags(A, np, gen);
            ok = !(nfl & PFF_INOC) && pf_free(A, np, nfl);
            if (ok) pf_set(A, np, gen, nfl | PFF_INOC, dl);
        }
        unsigned okm = __ballot_sync(FULLM, ok);
        int slot = oi + __popc(okm & below); // queue order = direction order
        if (ok) qpos[slot] = (uint16_t)np;
        oi += __popc(okm);
        __syncwarp();
    }
    __syncwarp();
    return result;
}

// GreedyPathFinding.findPathToPositionInRange (GreedyPathFinding.java:53-84): the free neighbour with the smallest squared
// distance to the target -- the first free direction is always taken, a later one only if strictly closer.  "Already in
// range" compares the SQUARED distance with the unsquared range (:66), as the reference does.  Lanes 0..3 look at one
// neighbour each; the choice is made on warp-uniform values.
DEVN int pf_greedy(Game &g, int s, int tx, int ty, int range, int nd) {
    uint32_t sw = g.w0()[s];
    int sx = u_x(sw), sy = u_y(sw), start = cell_of(g, sw);
    if (range >= 0 && (tx - sx) * (tx - sx) + (ty - sy) * (ty - sy) <= range) return -1; // range < 0: findPath (:16-47) has no such test
    int dl = g.lane & 3, np = start + doff(g, dl);
    bool fre = g.grid()[np] == 0 && g.resv()[np] == 0;
#pragma unroll 1
    for (int k = 0; k < nd && fre; k++) { // ru.getPositionsUsed(): the target cells of the desires chosen so far
        uint32_t A0 = g.pa0()[k];
        if (a_uses_cell(a_type(A0)) && linear_target_cell(g, g.w0()[g.pslot()[k]], g.pa1()[k]) == np) fre = false;
    }
    int x = sx + ddx(dl), y = sy + ddy(dl);
    int d = fre ? (tx - x) * (tx - x) + (ty - y) * (ty - y) : -1;
    int dir = -1, min_d = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int di = __shfl_sync(FULLM, d, i);
        if (di >= 0 && (dir == -1 || di < min_d)) { min_d = di; dir = i; }
    }
    return dir;
}

// FloodFillPathFinding.findPathToPositionInRange (FloodFillPathFinding.java:168-213) with its per-instance cache of distance maps
// (:17, one int[w][h] per target position, kept across cycles and queried again as long as the step it suggests is free).  The
// instance is the AI's: a game carries one cache per player in HBM -- [last frame, valid bits per target position, W*H maps of W*H
// u16 distances (0xFFFF = Integer.MAX_VALUE)] -- allocated when a player's pathfinder is set to MRTS_PF_FLOODFILL.
// doFloodFill (:47-107): breadth first from the TARGET over cells that are free (gs.getAllFree: no wall, no unit, not the target of an
// in-flight MOVE / PRODUCE) and not used by this cycle's earlier desires; neighbours in the order left, up, right, down; it stops after
// the node whose neighbourhood contains the start cell.  getAction (:139-166): the neighbour of the start with the smallest distance,
// first minimum in the order left, up, right, down.
#ifndef MRTS_TU_RUSH_ONLY
DEV int ff_get_action(const Game &g, const uint16_t *dist, int x, int y) {
    const int W = g.W, H = g.H;
    int d0 = x > 0 ? dist[(x - 1) + y * W] : 0xFFFF, d1 = y > 0 ? dist[x + (y - 1) * W] : 0xFFFF;
    int d2 = x + 1 < W ? dist[(x + 1) + y * W] : 0xFFFF, d3 = y + 1 < H ? dist[x + (y + 1) * W] : 0xFFFF;
    int index = 0, mn = d0;
    if (d1 < mn) { index = 1; mn = d1; }
    if (d2 < mn) { index = 2; mn = d2; }
    if (d3 < mn) { index = 3; mn = d3; }
    if (mn == 0xFFFF) return -1;
    return index == 0 ? 3 : index - 1; // left, up, right, down -> DIRECTION_LEFT (3), UP (0), RIGHT (1), DOWN (2)
}
DEVN int pf_floodfill(Game &g, int player, int s, int tx, int ty, int range, int nd) {
    const PfArr A = g.as_sm ? pf_arrays<true>(g) : pf_arrays<false>(g);
    const int lane = g.lane, W = g.W, H = g.H, cells = W * H;
    uint32_t sw = g.w0()[s];
    const int sx = u_x(sw), sy = u_y(sw);
    if (tx < 0 || ty < 0 || tx >= W || ty >= H) return -1; // the reference indexes its distance array with the target (an exception off the map): null
    if (range < 0) range = 0;
    if ((sx - tx) * (sx - tx) + (sy - ty) * (sy - ty) <= range * range) return -1; // already there
    if (!g.ff_cache) return -1;
    int32_t *hdr = (int32_t *)(g.ff_cache + (size_t)player * g.ff_stride);       // [0] lastFrame, [1..] valid bits
    uint32_t *valid = (uint32_t *)(hdr + 1);
    uint16_t *maps = (uint16_t *)(hdr + 1 + ((cells + 31) >> 5) + (((cells + 31) >> 5) & 1 ? 0 : 1)); // 8-byte aligned start
    const int time = g.hdr()[H_TIME], target = tx + ty * W;
    __syncwarp();
    if (time < hdr[0]) { // a new game: cache.clear()
        __syncwarp();
        for (int i = lane; i < ((cells + 31) >> 5); i += 32) valid[i] = 0;
    }
    __syncwarp();
    if (lane == 0) hdr[0] = time;
    // initFree: the cells used by the desires chosen so far this cycle
    int gen = *A.gen + 1;
    __syncwarp();
    if (gen >= PF_GEN_LIMIT) { int pcells = A.P * (H + 2); for (int i = lane; i < pcells; i += 32) A.mark[i] = 0; gen = 1; }
    if (lane == 0) *A.gen = (uint16_t)gen;
    __syncwarp();
    for (int k = lane; k < nd; k += 32) {
        uint32_t A0 = g.pa0()[k];
        if (a_uses_cell(a_type(A0))) { int pc = linear_target_cell(g, g.w0()[g.pslot()[k]], g.pa1()[k]); if (pc >= 0) pf_set(A, pc, gen, PFF_BLOCKED); }
    }
    __syncwarp();
    uint16_t *dist = maps + (size_t)target * cells;
    bool cached = (valid[target >> 5] >> (target & 31)) & 1;
    __syncwarp();
    if (cached) {
        int dir = ff_get_action(g, dist, sx, sy);
        if (dir >= 0) {
            int pc = cell_of(g, sw) + doff(g, dir);
            bool ok = !(pf_flags(A, pc, gen) & PFF_BLOCKED) && g.grid()[pc] == 0 && g.resv()[pc] == 0; // free[][] and gs.free()
            if (ok) return dir;
        }
        // the suggested step is taken, or there is no step: cache.remove(targetpos) and compute again
    }
    // calculateDistances (:108-126)
    __syncwarp();
    for (int i = lane; i < cells; i += 32) dist[i] = 0xFFFF;
    __syncwarp();
    uint16_t *q = A.next; // the fringe: linear positions x + y * W
    int qn = 1, index = 0;
    if (lane == 0) { dist[target] = 0; q[0] = (uint16_t)target; }
    __syncwarp();
    const int dl = lane & 3; // this lane's neighbour: 0 left, 1 up, 2 right, 3 down
    const int ddxl = dl == 0 ? -1 : (dl == 2 ? 1 : 0), ddyl = dl == 1 ? -1 : (dl == 3 ? 1 : 0);
#pragma unroll 1
    while (index < qn) {
        int cur = q[index], x = cur % W, y = cur / W, dcur = dist[cur];
        int nx = x + ddxl, ny = y + ddyl;
        bool inb = nx >= 0 && ny >= 0 && nx < W && ny < H, ok = false, hit = false;
        if (lane < 4) {
            hit = nx == sx && ny == sy;
            if (inb) {
                int pc = (ny + 1) * A.P + nx + 1;
                ok = dist[nx + ny * W] == 0xFFFF && !(pf_flags(A, pc, gen) & PFF_BLOCKED) && g.grid()[pc] == 0 && g.resv()[pc] == 0;
            }
        }
        unsigned okm = __ballot_sync(FULLM, ok), hitm = __ballot_sync(FULLM, hit);
        if (ok) { int slot = qn + __popc(okm & ((1u << lane) - 1)); dist[nx + ny * W] = (uint16_t)(dcur + 1); q[slot] = (uint16_t)(nx + ny * W); }
        qn += __popc(okm);
        __syncwarp();
        if (hitm) break;
        index++;
    }
    __syncwarp();
    if (lane == 0) valid[target >> 5] |= 1u << (target & 31);
    __syncwarp();
    return ff_get_action(g, dist, sx, sy);
}
#endif

// range < 0: PathFinding.findPath (A* and BFS: findPathToPositionInRange with range 0)
// (A* and BFS pop the start first and stop there when it is within range: null.  Harvesters next to their resource or base ask
// exactly that, half of their queries, so it is answered before the search is set up.)
DEV bool pf_start_in_range(const Game &g, int s, int tx, int ty, int range) {
    uint32_t sw = g.w0()[s];
    int dx = u_x(sw) - tx, dy = u_y(sw) - ty;
    return dx * dx + dy * dy <= range * range;
}
DEV int pf_find(Game &g, int kind, int s, int tx, int ty, int range, int nd) {
#ifdef MRTS_TU_RUSH_ONLY
    if (range < 0) range = 0;
    if (pf_start_in_range(g, s, tx, ty, range)) return -1;
    return pf_find_t<true, 0>(g, 0, s, tx, ty, range, nd); // the lean copy: A*, scratch in shared memory
#else
    if (kind == 2) return pf_greedy(g, s, tx, ty, range, nd);
    if (kind == 3) return pf_floodfill(g, u_pl(g.w0()[s]) == 2 ? 1 : 0, s, tx, ty, range, nd);
    if (range < 0) range = 0;
    if (pf_start_in_range(g, s, tx, ty, range)) return -1;
    return g.as_sm ? pf_find_t<true>(g, kind, s, tx, ty, range, nd) : pf_find_t<false>(g, kind, s, tx, ty, range, nd);
#endif
}

// ---- AbstractionLayerAI -------------------------------------------------------------------------------------------------
struct ScriptCtx { int player, pf, par0, par1, nd; };

// GameState.isUnitActionAllowed (GameState.java:434-457)
DEV bool unit_action_allowed(const Game &g, const ScriptCtx &c, int s, uint32_t A0, int A1) {
    uint32_t w = g.w0()[s];
    int at = a_type(A0), pc = cell_of(g, w), tc = -1, cost = 0;
    if (at == ACT_MOVE) {
        tc = pc + doff(g, A1);
        if (g.grid()[tc] != 0) return false; // out of bounds, wall or occupied
    } else if (at == ACT_PRODUCE) { tc = target_cell(g, pc, A1); cost = ut_cost(g, a_utype(A0)); }
    if (tc >= 0 && g.resv()[tc] != 0) return false;
    int pl = u_pl(w);
    int s0 = (pl == 1 ? cost : 0) + c.par0, s1 = (pl == 2 ? cost : 0) + c.par1;
    if (c.par0 != 0 && s0 > 0 && s0 > g.hdr()[H_RES0]) return false;
    if (c.par1 != 0 && s1 > 0 && s1 > g.hdr()[H_RES1]) return false;
    return true;
}

DEV bool mk_move(const Game &g, const ScriptCtx &c, int s, int dir, uint32_t &A0, int &A1) {
    if (dir < 0) return false;
    A0 = ACT_MOVE | A0_NOUT; A1 = dir;
    return true;
}

// Train.score (Train.java:98-126): minus the Manhattan distance to the closest resource (harvesters) / enemy unit (others)
DEV int train_score(const Game &g, int x, int y, int type, int pl) {
    int n = g.hdr()[H_NUNITS], distance = 0;
    bool harvester = (ut_flags(g, type) & UF_HARVEST) != 0;
    w_argmin(g, n, [&](int, uint32_t w) {
        bool ok = harvester ? (ut_flags(g, u_type(w)) & UF_RESOURCE) != 0 : (u_pl(w) != 0 && u_pl(w) != pl);
        return ok ? iabs(u_x(w) - x) + iabs(u_y(w) - y) : -1;
    }, &distance);
    return -distance;
}

DEV bool cell_gs_free(const Game &g, int pc) { return g.grid()[pc] == 0 && g.resv()[pc] == 0; }


#ifndef MRTS_TU_RUSH_ONLY
// findPathToPositionInRange with the target as the reference passes it, a LINEAR position x + y * width that the pathfinders split
// again with Java's % and / (a column off the map aliases into the neighbouring row, a negative position keeps a negative x)
DEV int pf_find_lin(Game &g, int kind, int s, int pos, int range, int nd) {
    return pf_find(g, kind, s, pos % g.W, pos / g.W, range, nd);
}
// PhysicalGameState.getUnitAt: the unit's slot, or -1 (null) for a free cell, a wall or a position off the map; |x|, |y| at most one off
DEV int unit_at(const Game &g, int x, int y) { int v = g.grid()[(y + 1) * g.P + x + 1]; return (v == 0 || v == 0xFF) ? -1 : v - 1; }

// cRush/CRanged_Tactic.java:77-307 (execute, squareMove).  Distances are square roots of integers compared with each other or
// with integers: their squares are compared instead.  Whole warp, uniform.
DEVN bool tactic_execute(Game &g, const ScriptCtx &c, int s, uint32_t &A0, int &A1) {
    const uint32_t X0 = g.x0()[s], X1 = g.x1()[s], w = g.w0()[s], tw = g.w0()[aa_target(X1) - 1];
    const int n = g.hdr()[H_NUNITS], W = g.W;
    const int x = u_x(w), y = u_y(w), t = u_type(w), pl = u_pl(w), tx = u_x(tw), ty = u_y(tw);
    int hx = aa_thx(X1), hy = aa_thy(X0), ex = aa_tex(X0), ey = aa_tey(X0);
    if (hx == 0xFF) { hx = x; hy = y; }   // home == null: the unit itself (:84-86)
    if (ex == 0xFF) { ex = tx; ey = ty; } // enemyBase == null: the target (:88-90)
    const int range = ut_range(g, t);
    const int rd2 = (hx - x) * (hx - x) + (hy - y) * (hy - y), d2 = (tx - x) * (tx - x) + (ty - y) * (ty - y);
    // `u2.getPlayer() != p.getID()`: neutral units count as well, only their types never match
    const int n_enemy_bases = w_count(g, n, [&](int, uint32_t ow) { return u_pl(ow) != pl && u_type(ow) == UT_BASE; });
    const int enemy_attack_units = w_count(g, n, [&](int, uint32_t ow) { int ot = u_type(ow); return u_pl(ow) != pl && (ot == UT_RANGED || ot == UT_HEAVY || ot == UT_LIGHT); });
    const int enemy_workers = w_count(g, n, [&](int, uint32_t ow) { return u_pl(ow) != pl && u_type(ow) == UT_WORKER; });
    const int cutoff = (g.W * g.H > 3000) ? 15000 : 5000;
    const bool time_to_attack = ((enemy_workers < 2 * n_enemy_bases || n_enemy_bases == 0) && enemy_attack_units == 0) || g.hdr()[H_TIME] > cutoff;
    // nearestRangedAlly(enemyBase) :367-389: the first own Ranged unit with the smallest distance to the enemy base
    const int ally = w_argmin(g, n, [&](int, uint32_t ow) { return (u_pl(ow) == pl && u_type(ow) == UT_RANGED) ? (u_x(ow) - ex) * (u_x(ow) - ex) + (u_y(ow) - ey) * (u_y(ow) - ey) : -1; });
    int ax = 0, ay = 0, ad2 = 0;
    if (ally >= 0) { uint32_t aw = g.w0()[ally]; ax = u_x(aw); ay = u_y(aw); ad2 = (tx - ax) * (tx - ax) + (ty - ay) * (ty - ay); }
    const int tpos = tx + ty * W;
    int dir = -1;
    if (d2 <= range * range) { A0 = ACT_ATTACK | A0_NOUT | ((uint32_t)tx << 16) | ((uint32_t)ty << 24); A1 = -1; return true; } // :152, :179
    if (t == UT_WORKER) { // :149-176
        if (time_to_attack || ally < 0) dir = pf_find_lin(g, c.pf, s, tpos, range, c.nd);
        else {
            dir = pf_find_lin(g, c.pf, s, d2 > ad2 ? tpos : ax + ay * W, range, c.nd);
            if (dir < 0) dir = pf_find_lin(g, c.pf, s, (ax - 1) + ay * W, range + 1, c.nd);
            if (dir < 0) dir = pf_find_lin(g, c.pf, s, tpos, range, c.nd);
        }
    } else if (ally < 0 || ally == s) { // the unit leads :182-194
        if (time_to_attack && u_type(tw) == UT_BASE) dir = pf_find_lin(g, c.pf, s, tpos, range, c.nd);
        else if (rd2 < 25 || (ex - x) * (ex - x) + (ey - y) * (ey - y) > (ex - hx) * (ex - hx) + (ey - hy) * (ey - hy)) dir = pf_find_lin(g, c.pf, s, ex + ey * W, range, c.nd);
    } else if (time_to_attack) { // :195-216 (d > range here, so of RangedAttack's three cases only the approach is left)
        dir = pf_find_lin(g, c.pf, s, tpos, range, c.nd);
    } else { // line up next to the leading ranged unit :218-268, squareMove :290-343
        auto eb2 = [&](int px, int py) { return (ex - px) * (ex - px) + (ey - py) * (ey - py); };
        const int sgn = eb2(ax, ay + 1) > eb2(ax, ay) ? 1 : -1; // below / right of the leader, or above / left
#pragma unroll 1
        for (;;) {
            int a = unit_at(g, ax, ay + sgn), b = unit_at(g, ax + sgn, ay + sgn), cc = unit_at(g, ax + sgn, ay);
            bool found = true;
            if (a >= 0 && a != s && b >= 0 && b != s && cc >= 0 && cc != s) { uint32_t cw = g.w0()[cc]; ax = u_x(cw); ay = u_y(cw); found = false; }
            a = unit_at(g, ax, ay + sgn); b = unit_at(g, ax + sgn, ay + sgn); cc = unit_at(g, ax + sgn, ay);
            if (found || a < 0 || b < 0 || cc < 0) break;
        }
        const int sg = eb2(ax, ay + 1) > eb2(ax, ay) ? 1 : -1; // squareMove decides again, around the unit it was handed
        const int a = unit_at(g, ax, ay + sg), b = unit_at(g, ax + sg, ay + sg), cc = unit_at(g, ax + sg, ay);
        if (s == a || s == b || s == cc) return false;
        // pf.findPath = range 0.  The reference computes all six paths and uses one: only that one is searched here
        if (a < 0) dir = pf_find_lin(g, c.pf, s, ax + (ay + sg) * W, -1, c.nd);
        else if (cc < 0) dir = pf_find_lin(g, c.pf, s, (ax + sg) + ay * W, -1, c.nd);
        else if (b < 0) dir = pf_find_lin(g, c.pf, s, (ax + sg) + (ay + sg) * W, -1, c.nd);
    }
    return mk_move(g, c, s, dir, A0, A1) && unit_action_allowed(g, c, s, A0, A1);
}
#endif

// AbstractAction.execute for slot s; returns true and (A0, A1) if it yields a unit action.  Whole warp, uniform.
DEVN bool aa_execute(Game &g, const ScriptCtx &c, int s, uint32_t &A0, int &A1) {
    uint32_t X0 = g.x0()[s], X1 = g.x1()[s], w = g.w0()[s];
    int x = u_x(w), y = u_y(w), t = u_type(w);
    switch (aa_kind(X0)) {
        case AA_ATTACK: { // Attack.java:51-64
            uint32_t tw = g.w0()[aa_target(X1) - 1];
            int dx = u_x(tw) - x, dy = u_y(tw) - y, range = ut_range(g, t);
            if (dx * dx + dy * dy <= range * range) { A0 = ACT_ATTACK | A0_NOUT | ((uint32_t)u_x(tw) << 16) | ((uint32_t)u_y(tw) << 24); A1 = -1; return true; }
            int dir = pf_find(g, c.pf, s, u_x(tw), u_y(tw), range, c.nd);
            if (mk_move(g, c, s, dir, A0, A1) && unit_action_allowed(g, c, s, A0, A1)) return true;
            return false;
        }
#ifndef MRTS_TU_RUSH_ONLY
        case AA_TACTIC: return tactic_execute(g, c, s, A0, A1);
        case AA_RANGED_ATTACK: { // cRush/RangedAttack.java:58-87: step back towards the barracks while a slower enemy is well inside
            // the range, shoot when in range, approach otherwise (the square roots compare like their integer squares)
            uint32_t tw = g.w0()[aa_target(X1) - 1];
            int dx = u_x(tw) - x, dy = u_y(tw) - y, range = ut_range(g, t), d2 = dx * dx + dy * dy, rd2 = 0;
            const int racks = aa_base(X1);
            uint32_t rw = 0;
            if (racks != (int)REF_NULL) { rw = g.w0()[racks - 1]; int rdx = u_x(rw) - x, rdy = u_y(rw) - y; rd2 = rdx * rdx + rdy * rdy; }
            const uint16_t *eta = (const uint16_t *)(g.utt() + MRTS_ETA_OFFSET);
            int dir;
            if (range >= 1 && d2 <= (range - 1) * (range - 1) && rd2 > 4 && eta[t * 8 + ACT_MOVE] < eta[u_type(tw) * 8 + ACT_MOVE])
                dir = pf_find(g, c.pf, s, u_x(rw), u_y(rw), range, c.nd);
            else if (d2 <= range * range) { A0 = ACT_ATTACK | A0_NOUT | ((uint32_t)u_x(tw) << 16) | ((uint32_t)u_y(tw) << 24); A1 = -1; return true; }
            else dir = pf_find(g, c.pf, s, u_x(tw), u_y(tw), range, c.nd);
            if (mk_move(g, c, s, dir, A0, A1) && unit_action_allowed(g, c, s, A0, A1)) return true;
            return false;
        }
#endif
        case AA_HARVEST: { // Harvest.java:72-113
            bool empty = u_res(g.w1()[s]) == 0;
            int other = empty ? aa_target(X1) : aa_base(X1);
            if (other == (int)REF_NULL) return false;
            uint32_t ow = g.w0()[other - 1];
            int dir = pf_find(g, c.pf, s, u_x(ow), u_y(ow), 1, c.nd);
            if (mk_move(g, c, s, dir, A0, A1)) return unit_action_allowed(g, c, s, A0, A1);
            int ox = u_x(ow), oy = u_y(ow), d = -1;
            if (ox == x && oy == y - 1) d = 0; else if (ox == x + 1 && oy == y) d = 1; else if (ox == x && oy == y + 1) d = 2; else if (ox == x - 1 && oy == y) d = 3;
            if (d < 0) return false;
            A0 = (empty ? ACT_HARVEST : ACT_RETURN) | A0_NOUT; A1 = d;
            return true;
        }
        case AA_BUILD: { // Build.java:54-77
            int bx = aa_bx(X0), by = aa_by(X0);
            int dir = pf_find(g, c.pf, s, bx, by, 1, c.nd);
            if (mk_move(g, c, s, dir, A0, A1)) return unit_action_allowed(g, c, s, A0, A1);
            int d = -1;
            if (bx == x && by == y - 1) d = 0;
            if (bx == x + 1 && by == y) d = 1;
            if (bx == x && by == y + 1) d = 2;
            if (bx == x - 1 && by == y) d = 3;
            if (d < 0) return false;
            A0 = ACT_PRODUCE | ((uint32_t)aa_type(X0) << 8); A1 = d;
            return unit_action_allowed(g, c, s, A0, A1);
        }
        case AA_MOVE: { // Move.java:49-55
            int dir = pf_find(g, c.pf, s, aa_bx(X0), aa_by(X0), -1, c.nd);
            if (mk_move(g, c, s, dir, A0, A1) && unit_action_allowed(g, c, s, A0, A1)) return true;
            return false;
        }
        case AA_TRAIN: { // Train.java:48-95
            int pc = cell_of(g, w), type = aa_type(X0), best_dir = -1, best = -1, pl = u_pl(w);
#pragma unroll 1
            for (int d = 0; d < 4; d++) {
                int nc = pc + doff(g, d);
                if (cell_gs_free(g, nc)) { int sc = train_score(g, x + ddx(d), y + ddy(d), type, pl); if (sc > best || best_dir == -1) { best = sc; best_dir = d; } }
            }
            __syncwarp();
            if (g.lane == 0) g.x0()[s] = X0 | 8u; // completed = true
            __syncwarp();
            if (best_dir != -1) { A0 = ACT_PRODUCE | ((uint32_t)type << 8); A1 = best_dir; return unit_action_allowed(g, c, s, A0, A1); }
            return false;
        }
    }
    return false;
}

// is the referenced unit in the unit list of the state the policy looks at?  (a unit hidden from a partially observable
// view is not: PartiallyObservableGameState removes it from its copy of the list)
DEV bool ref_alive(const Game &g, int r) { return r != (int)REF_NULL && r != (int)REF_DEAD && !(g.po_view && w_hidden(g.w0()[r - 1])); }

// AbstractAction.completed
DEV bool aa_completed(const Game &g, int s) {
    uint32_t X0 = g.x0()[s], X1 = g.x1()[s];
    switch (aa_kind(X0)) {
        case AA_TRAIN: return (X0 & 8u) != 0;
        case AA_BUILD: { int bx = aa_bx(X0), by = aa_by(X0); if (bx < 0 || bx >= g.W || by >= g.H) return false; int gv = g.grid()[(by + 1) * g.P + bx + 1]; return gv != 0 && gv != 0xFF; }
        case AA_HARVEST: return u_res(g.w1()[s]) > 0 ? !ref_alive(g, aa_base(X1)) : !ref_alive(g, aa_target(X1));
        case AA_ATTACK: case AA_RANGED_ATTACK: case AA_TACTIC: return !ref_alive(g, aa_target(X1)); // Attack.java:30-33, RangedAttack.java:36-39
        case AA_MOVE: { uint32_t w = g.w0()[s]; return u_x(w) == aa_bx(X0) && u_y(w) == aa_by(X0); } // Move.java:29-31
    }
    return true;
}

// AbstractionLayerAI.findBuildingPosition (:143-229) over PhysicalGameState.getAllFree (units and walls only)
DEVN int find_building_position(const Game &g, const int *reserved, int nres, int dX, int dY) {
    int W = g.W, H = g.H, maxl = H > W ? H : W;
#pragma unroll 1
    for (int l = 1; l < maxl; l++) {
#pragma unroll 1
        for (int side = 0; side < 4; side++) {
            int fx, fy, sx, sy; // first cell and step of the side's scan
            if (side == 0) { fy = dY - l; if (fy < 0) continue; fx = dX - l; sx = 1; sy = 0; }
            else if (side == 1) { fx = dX + l; if (fx >= W) continue; fy = dY - l; sx = 0; sy = 1; }
            else if (side == 2) { fy = dY + l; if (fy >= H) continue; fx = dX - l; sx = 1; sy = 0; }
            else { fx = dX - l; if (fx < 0) continue; fy = dY - l; sx = 0; sy = 1; }
#pragma unroll 1
            for (int k = 0; k <= 2 * l; k++) {
                int x = fx + k * sx, y = fy + k * sy;
                if (x < 0 || x >= W || y < 0 || y >= H) continue;
                int pos = x + y * W;
                bool rsv = false;
                for (int q = 0; q < nres; q++) rsv |= reserved[q] == pos;
                if (!rsv && g.grid()[(y + 1) * g.P + x + 1] == 0) return pos;
            }
        }
    }
    return -1;
}

// meleeUnitBehavior.  The rushes attack the closest enemy (WorkerRush.java:105-121, LightRush.java:141-159).  The defenses
// (WorkerDefense.java:117-146, LightDefense.java:142-165) do so only while that enemy, or the own base -- the LAST own base
// of the unit list, distance 0 without one -- is closer than height/2; otherwise they put an Attack with a null target,
// which translateActions finds completed and deletes (so the unit's next entry goes to the end of the map).
// (the defense and exploration variants are kept out of line: the generic kernel is instruction-cache bound, and the rushes
// of the benchmark configuration never run them)
DEVN void script_melee_special(Game &g, int s, int player, bool defense, bool explore, bool always) {
    int n = g.hdr()[H_NUNITS];
    uint32_t w = g.w0()[s];
    int cd = 0;
    int closest = w_argmin(g, n, [&](int, uint32_t ow) { return (u_pl(ow) != 0 && u_pl(ow) != player + 1) ? iabs(u_x(ow) - u_x(w)) + iabs(u_y(ow) - u_y(w)) : -1; }, &cd);
    if (defense) {
        int last = w_argmin(g, n, [&](int i, uint32_t ow) { return (u_pl(ow) == player + 1 && u_type(ow) == 1 /* baseType */) ? 255 - i : -1; });
        int mybase = 0;
        if (last >= 0) { uint32_t bw = g.w0()[last]; mybase = iabs(u_x(bw) - u_x(w)) + iabs(u_y(bw) - u_y(w)); }
        if (closest >= 0 && (always || cd < g.H / 2 || mybase < g.H / 2)) aa_put(g, s, player, AA_ATTACK, 0, 0, 0, closest + 1, REF_NULL); // always: WorkerRushPlusPlus.java:136-142
        else aa_put(g, s, player, AA_ATTACK, 0, 0, 0, REF_NULL, REF_NULL);
        return;
    }
    if (closest >= 0) aa_put(g, s, player, AA_ATTACK, 0, 0, 0, closest + 1, REF_NULL);
    else if (explore) {
        // PO*Rush.meleeUnitBehavior (POLightRush.java:56-77): no enemy in view, so walk to the nearest cell (squared distance,
        // first minimum in row-major order) that none of the player's units can see
        int ux = u_x(w), uy = u_y(w), cells = g.W * g.H;
        unsigned best = 0xFFFFFFFFu;
#pragma unroll 1
        for (int q = g.lane; q < cells; q += 32) {
            int y = q / g.W, x = q - y * g.W;
            if (!g.vis()[(y + 1) * g.P + x + 1]) { unsigned k = ((unsigned)((ux - x) * (ux - x) + (uy - y) * (uy - y)) << 14) | (unsigned)q; if (k < best) best = k; }
        }
        best = __reduce_min_sync(FULLM, best);
        if (best != 0xFFFFFFFFu) { int q = (int)(best & 0x3FFFu), y = q / g.W; aa_put(g, s, player, AA_MOVE, 0, q - y * g.W, y, REF_NULL, REF_NULL); }
    }
}
DEV void script_melee(Game &g, int s, int player, bool defense, bool explore = false, bool always = false) {
    if (defense || explore) { script_melee_special(g, s, player, defense, explore, always); return; }
    int n = g.hdr()[H_NUNITS];
    uint32_t w = g.w0()[s];
    int closest = w_argmin(g, n, [&](int, uint32_t ow) { return (u_pl(ow) != 0 && u_pl(ow) != player + 1) ? iabs(u_x(ow) - u_x(w)) + iabs(u_y(ow) - u_y(w)) : -1; });
    if (closest >= 0) aa_put(g, s, player, AA_ATTACK, 0, 0, 0, closest + 1, REF_NULL);
}

// harvest part of workersBehavior (WorkerRush.java:148-199, LightRush.java:203-252); true if the worker is still free
DEV bool script_harvest(Game &g, int s, int player, bool defense) {
    int n = g.hdr()[H_NUNITS];
    uint32_t w = g.w0()[s];
    int cres = w_argmin(g, n, [&](int, uint32_t ow) { return (ut_flags(g, u_type(ow)) & UF_RESOURCE) ? iabs(u_x(ow) - u_x(w)) + iabs(u_y(ow) - u_y(w)) : -1; });
    int cbase = w_argmin(g, n, [&](int, uint32_t ow) {
        return ((ut_flags(g, u_type(ow)) & UF_STOCKPILE) && u_pl(ow) == player + 1) ? iabs(u_x(ow) - u_x(w)) + iabs(u_y(ow) - u_y(w)) : -1; });
    uint32_t X0 = g.x0()[s], X1 = g.x1()[s];
    bool is_h = aa_kind(X0) == AA_HARVEST;
    if (defense) { // WorkerDefense.java:197-209, LightDefense.java:236-244: no special case for a worker carrying resources,
        // and a worker that cannot harvest is never sent to attack
        if (cres >= 0 && cbase >= 0 && (!is_h || aa_target(X1) != cres + 1 || aa_base(X1) != cbase + 1)) aa_put(g, s, player, AA_HARVEST, 0, 0, 0, cres + 1, cbase + 1);
        return false;
    }
    if (u_res(g.w1()[s]) > 0) {
        if (cbase >= 0) {
            if (!is_h || aa_base(X1) != cbase + 1) aa_put(g, s, player, AA_HARVEST, 0, 0, 0, REF_NULL, cbase + 1);
            return false;
        }
    } else if (cres >= 0 && cbase >= 0) {
        if (!is_h || aa_target(X1) != cres + 1 || aa_base(X1) != cbase + 1) aa_put(g, s, player, AA_HARVEST, 0, 0, 0, cres + 1, cbase + 1);
        return false;
    }
    return true;
}

// desired position: the builder's own cell unless (dX, dY) is given
DEV void script_build_if_not(Game &g, int s, int player, int type, int *reserved, int &nres, int dX = -1000, int dY = 0) { // buildIfNotAlreadyBuilding :231-245
    uint32_t X0 = g.x0()[s];
    if (!(aa_kind(X0) == AA_BUILD && aa_type(X0) == type)) {
        uint32_t w = g.w0()[s];
        if (dX == -1000) { dX = u_x(w); dY = u_y(w); }
        int pos = find_building_position(g, reserved, nres, dX, dY);
        int bx = pos < 0 ? -1 : pos % g.W, by = pos < 0 ? 0 : pos / g.W; // Java: -1 % w == -1, -1 / w == 0
        aa_put(g, s, player, AA_BUILD, type, bx, by, REF_NULL, REF_NULL);
        if (nres < 4) reserved[nres++] = pos;
    }
}


// translateActions (AbstractionLayerAI.java:58-113): the player's abstract actions in insertion order become desires, the
// desires a player action (appended to the pending list from pn; returns the new pending count, uniform)
DEV int translate_actions(Game &g, int player, int pathfinder, int par0, int par1, int pn) {
    const int n = g.hdr()[H_NUNITS], pl = player + 1, pres = g.hdr()[H_RES0 + player];
    ScriptCtx c; c.player = player; c.pf = pathfinder; c.par0 = par0; c.par1 = par1; c.nd = 0;
    int nd = 0;
    // The reference walks the player's map entries in insertion order: a completed entry is dropped, an entry of an idle unit is
    // executed.  completed() reads nothing an execute() changes (the Train flag belongs to the executed entry itself), so all
    // entries are tested at once, one per lane, and the completed ones dropped on the spot; only the few entries left to EXECUTE
    // are put into insertion order (an insertion sort by sequence number on lane 0: slot order is nearly that order already).
    int ne = 0;
    __syncwarp();
#pragma unroll 1
    for (int base = 0; base < n; base += 32) {
        int i = base + g.lane;
        bool has = i < n && aa_kind(g.x0()[i]) != AA_NONE && u_pl(g.w0()[i]) == pl;
        bool comp = has && aa_completed(g, i);
        if (comp) g.x0()[i] &= ~7u; // toDelete (a dead unit's entry vanished with its slot)
        bool run = has && !comp && a_type(g.a0()[i]) == AT_IDLE;
        unsigned m = __ballot_sync(FULLM, run);
        if (run) g.list()[ne + __popc(m & ((1u << g.lane) - 1))] = (uint8_t)i;
        ne += __popc(m);
    }
    __syncwarp();
    if (g.lane == 0) {
#pragma unroll 1
        for (int k = 1; k < ne; k++) {
            int i = g.list()[k];
            uint32_t q = aa_seq(g.x0()[i], g.x1()[i]);
            int j = k;
#pragma unroll 1
            while (j > 0) { int o = g.list()[j - 1]; if (aa_seq(g.x0()[o], g.x1()[o]) <= q) break; g.list()[j] = (uint8_t)o; j--; }
            g.list()[j] = (uint8_t)i;
        }
    }
    __syncwarp();
#pragma unroll 1
    for (int r = 0; r < ne; r++) {
        int best = g.list()[r];
        uint32_t A0; int A1;
        c.nd = nd;
        // Desires are staged behind the part of the pending list that is already final ([0, pn): player 0's list when this is
        // player 1) and pf_find reads them as [0, nd) through a shifted window.
        int pv = g.pview;
        g.pview = pv + pn;
        bool got = aa_execute(g, c, best, A0, A1);
        g.pview = pv;
        __syncwarp();
        if (got) {
            if (g.lane == 0) { g.pslot()[pn + nd] = (uint8_t)best; g.pa0()[pn + nd] = A0; g.pa1()[pn + nd] = A1; }
            nd++;
        }
        __syncwarp();
    }
    int out = pn;
    if (g.lane == 0) {
        // compose desires against gs.getResourceUsage() (:93-101): pa.consistentWith(r2) with pa.r = in-flight + accepted
        int acc0 = par0, acc1 = par1, m = pn;
#pragma unroll 1
        for (int k = 0; k < nd; k++) {
            int s = g.pslot()[pn + k]; uint32_t A0 = g.pa0()[pn + k]; int A1 = g.pa1()[pn + k];
            uint32_t w = g.w0()[s];
            int at = a_type(A0), tc = -1, cost = 0;
            if (a_uses_cell(at)) { tc = target_cell(g, cell_of(g, w), A1); if (at == ACT_PRODUCE) cost = ut_cost(g, a_utype(A0)); }
            bool ok = !(tc >= 0 && (g.resv()[tc] != 0 || ((g.claim()[tc] >> player) & 1)));
            if (ok && cost != 0) { int tot = (pl == 1 ? acc0 : acc1) + cost; if (tot > 0 && tot > pres) ok = false; }
            if (ok) {
                if (tc >= 0) g.claim()[tc] |= (uint8_t)(1 << player);
                if (pl == 1) acc0 += cost; else acc1 += cost;
                g.pslot()[m] = (uint8_t)s; g.pa0()[m] = A0; g.pa1()[m] = A1; m++;
            }
        }
        // the claims only modelled pa.r while composing; drop them (a desire may still be replaced by NONE in issueSafe)
#pragma unroll 1
        for (int q = pn; q < m; q++) {
            uint32_t A0 = g.pa0()[q];
            if (a_uses_cell(a_type(A0))) g.claim()[target_cell(g, cell_of(g, g.w0()[g.pslot()[q]]), g.pa1()[q])] = 0;
        }
        out = m;
    }
    __syncwarp();
    out = __shfl_sync(FULLM, out, 0);
    { // PlayerAction.fillWithNones(gs, player, 10) (PlayerAction.java:217-235): the player's idle units without an action, in
      // unit-list order; all lanes test 32 units at a time and append by ballot
        const int m0 = out;
#pragma unroll 1
        for (int base = 0; base < n; base += 32) {
            int i = base + g.lane;
            bool need = i < n && u_pl(g.w0()[i]) == pl && a_type(g.a0()[i]) == AT_IDLE;
            if (need) {
#pragma unroll 1
                for (int q = pn; q < m0; q++) if (g.pslot()[q] == i) { need = false; break; }
            }
            unsigned mk = __ballot_sync(FULLM, need);
            if (need) { int q = out + __popc(mk & ((1u << g.lane) - 1)); g.pslot()[q] = (uint8_t)i; g.pa0()[q] = ACT_NONE | A0_NOUT; g.pa1()[q] = 10; }
            out += __popc(mk);
        }
    }
    __syncwarp();
    return out;
}

#ifndef MRTS_TU_RUSH_ONLY
// CRush_V1.rangedUnitBehavior (cRush/CRush_V1.java:194-220): closest enemy and closest own barracks in ONE pass over the unit
// list that shares its running distance between the two searches -- order-dependent, so every lane walks the list
DEVN void script_crush_ranged(Game &g, int s, int player) {
    const int n = g.hdr()[H_NUNITS], pl = player + 1;
    const uint32_t w = g.w0()[s];
    int enemy = -1, racks = -1, cd = 0;
#pragma unroll 1
    for (int i = 0; i < n; i++) {
        uint32_t ow = g.w0()[i];
        bool en = u_pl(ow) != 0 && u_pl(ow) != pl, rk = u_type(ow) == UT_BARRACKS && u_pl(ow) == pl;
        if (en || rk) {
            int d = iabs(u_x(ow) - u_x(w)) + iabs(u_y(ow) - u_y(w));
            if (en && (enemy < 0 || d < cd)) { enemy = i; cd = d; }
            if (rk && (racks < 0 || d < cd)) { racks = i; cd = d; }
        }
    }
    if (enemy >= 0) aa_put(g, s, player, AA_RANGED_ATTACK, 0, 0, 0, enemy + 1, racks < 0 ? (int)REF_NULL : racks + 1);
}

// CRush_V2.meleeUnitBehavior (cRush/CRush_V2.java:174-219) / rangedUnitBehavior (:221-264): closest enemy, own barracks, own base
// and enemy base in one pass with one shared running distance (an enemy base is tested twice: as an enemy, then as a base).
// Before cycle 400 and on small maps everything but the Ranged units simply attacks; otherwise CRanged_Tactic.
DEVN void script_crush2_combat(Game &g, int s, int player, bool ranged, bool rush) {
    const int n = g.hdr()[H_NUNITS], pl = player + 1;
    const uint32_t w = g.w0()[s];
    int enemy = -1, racks = -1, base = -1, ebase = -1, cd = 0;
#pragma unroll 1
    for (int i = 0; i < n; i++) {
        uint32_t ow = g.w0()[i];
        int op = u_pl(ow), ot = u_type(ow);
        bool en = op != 0 && op != pl, own_b = op == pl && (ot == UT_BASE || ot == UT_BARRACKS), eb = ot == UT_BASE && op != pl;
        if (en || own_b || eb) {
            int d = iabs(u_x(ow) - u_x(w)) + iabs(u_y(ow) - u_y(w));
            if (en && (enemy < 0 || d < cd)) { enemy = i; cd = d; }
            if (own_b && ot == UT_BARRACKS && (racks < 0 || d < cd)) { racks = i; cd = d; }
            if (own_b && ot == UT_BASE && (base < 0 || d < cd)) { base = i; cd = d; }
            if (eb && (ebase < 0 || d < cd)) { ebase = i; cd = d; }
        }
    }
    if (enemy < 0) return;
    if (!ranged && (g.hdr()[H_TIME] < 400 || rush)) { aa_put(g, s, player, AA_ATTACK, 0, 0, 0, enemy + 1, REF_NULL); return; }
    int hx = 0xFF, hy = 0, ex = 0xFF, ey = 0;
    if (base >= 0) { uint32_t bw = g.w0()[base]; hx = u_x(bw); hy = u_y(bw); }
    if (ebase >= 0) { uint32_t bw = g.w0()[ebase]; ex = u_x(bw); ey = u_y(bw); }
    aa_put(g, s, player, AA_TACTIC, hy, ex, ey, enemy + 1, hx);
}

// CRush_V2.workersBehavior :335-383: a free worker leaves a resource alone that lies closer to the enemy base than to its own
// base (distance() is 0.0 when either end is null: without an enemy base nobody harvests)
DEVN void script_crush2_harvest(Game &g, int s, int player) {
    const int n = g.hdr()[H_NUNITS], pl = player + 1;
    const uint32_t w = g.w0()[s];
    int cres = -1, ebase = -1, cd = 0;
#pragma unroll 1
    for (int i = 0; i < n; i++) {
        uint32_t ow = g.w0()[i];
        bool rs = (ut_flags(g, u_type(ow)) & UF_RESOURCE) != 0, eb = u_type(ow) == UT_BASE && u_pl(ow) != pl;
        if (rs || eb) {
            int d = iabs(u_x(ow) - u_x(w)) + iabs(u_y(ow) - u_y(w));
            if (rs && (cres < 0 || d < cd)) { cres = i; cd = d; }
            if (eb && (ebase < 0 || d < cd)) { ebase = i; cd = d; }
        }
    }
    int cbase = w_argmin(g, n, [&](int, uint32_t ow) {
        return ((ut_flags(g, u_type(ow)) & UF_STOCKPILE) && u_pl(ow) == pl) ? iabs(u_x(ow) - u_x(w)) + iabs(u_y(ow) - u_y(w)) : -1; });
    if (cres < 0) return;
    const uint32_t rw = g.w0()[cres];
    int de = 0, db = 0;
    if (ebase >= 0) { uint32_t bw = g.w0()[ebase]; de = (u_x(bw) - u_x(rw)) * (u_x(bw) - u_x(rw)) + (u_y(bw) - u_y(rw)) * (u_y(bw) - u_y(rw)); }
    if (cbase >= 0) { uint32_t bw = g.w0()[cbase]; db = (u_x(bw) - u_x(rw)) * (u_x(bw) - u_x(rw)) + (u_y(bw) - u_y(rw)) * (u_y(bw) - u_y(rw)); }
    if (de < db || cbase < 0) return;
    uint32_t X0 = g.x0()[s], X1 = g.x1()[s];
    if (!(aa_kind(X0) == AA_HARVEST) || aa_target(X1) != cres + 1 || aa_base(X1) != cbase + 1) aa_put(g, s, player, AA_HARVEST, 0, 0, 0, cres + 1, cbase + 1);
}

// CRush_V1.getAction (cRush/CRush_V1.java:68-131) followed by translateActions.  On maps of at most 144 cells: a worker rush
// that keeps one harvester per base (rushWorkersBehavior :329-416, rushBaseBehavior :324-326).  On larger maps: nbases + 1
// harvesters, one barracks training Ranged units (workersBehavior :222-321, baseBehavior :133-168, barracksBehavior :170-174),
// every other worker fights.  The AI's `buildingRacks` field is a bit of header word H_AIFLAGS; `resourcesUsed` is written
// and read within one getAction.
// v2: CRush_V2.java -- the base also trains a worker when there are more than 6 Ranged units (:154), combat units follow
// CRanged_Tactic, free workers weigh resources against the enemy base.
DEVN int policy_crush(Game &g, int player, int pathfinder, int pn, bool v2) {
    int par0, par1;
    reserved_resources(g, par0, par1);
    __syncwarp();
    const int n = g.hdr()[H_NUNITS], pl = player + 1, pres = g.hdr()[H_RES0 + player];
    const bool rush = g.W * g.H <= 144;
    auto own_harvester = [&](int, uint32_t w) { return u_pl(w) == pl && (ut_flags(g, u_type(w)) & UF_HARVEST) != 0; };
    const int nbases = w_count(g, n, [&](int, uint32_t w) { return u_pl(w) == pl && u_type(w) == UT_BASE; });
    const int nbarracks = w_count(g, n, [&](int, uint32_t w) { return u_pl(w) == pl && u_type(w) == UT_BARRACKS; });
    const int nworkers = w_count(g, n, [&](int, uint32_t w) { return u_pl(w) == pl && u_type(w) == UT_WORKER; });
    const int nw = w_count(g, n, own_harvester);
    const int nranged = v2 ? w_count(g, n, [&](int, uint32_t w) { return u_pl(w) == pl && u_type(w) == UT_RANGED; }) : 0;
    bool building = ((g.hdr()[H_AIFLAGS] >> player) & 1) != 0;
    const bool was_building = building;
    int resources_used = 0;
    if (nw > 0) {
        // freeWorkers = the first nfree own harvesters of the unit list (minus the `taken` builders), battleWorkers = the rest
        const int keep = rush ? nbases : nbases + 1;
        const int nfree = (rush && pres == 0) ? 0 : (nw > keep ? keep : nw);
        int reserved[4] = {0, 0, 0, 0}, nres = 0, taken = 0, wi = -1;
        if (nbases == 0 && taken < nfree && pres >= ut_cost(g, UT_BASE)) { wi = w_next(g, n, wi, own_harvester); taken++; script_build_if_not(g, wi, player, UT_BASE, reserved, nres); }
        if (!rush) {
            if (nbarracks == 0 && taken < nfree && nworkers > 1 && pres >= ut_cost(g, UT_BARRACKS)) {
                wi = w_next(g, n, wi, own_harvester); taken++; script_build_if_not(g, wi, player, UT_BARRACKS, reserved, nres);
                resources_used += ut_cost(g, UT_BARRACKS);
                building = true;
            } else resources_used = ut_cost(g, UT_BARRACKS) * nbarracks;
            if (nbarracks > 1) building = true;
        }
        int ord = 0;
#pragma unroll 1
        for (int i = w_next(g, n, -1, own_harvester); i >= 0; i = w_next(g, n, i, own_harvester), ord++)
            if (ord >= nfree) { if (v2) script_crush2_combat(g, i, player, false, rush); else script_melee(g, i, player, false); }
        ord = 0;
#pragma unroll 1
        for (int i = w_next(g, n, -1, own_harvester); i >= 0 && ord < nfree; i = w_next(g, n, i, own_harvester), ord++)
            if (ord >= taken) { if (v2 && !rush) script_crush2_harvest(g, i, player); else script_harvest(g, i, player, true); }
    }
    if (building != was_building) { __syncwarp(); if (g.lane == 0) g.hdr()[H_AIFLAGS] |= 1 << player; __syncwarp(); }
    w_for_each(g, n, [&](int i, uint32_t w) { return u_type(w) == UT_BASE && u_pl(w) == pl && a_type(g.a0()[i]) == AT_IDLE; }, [&](int i) {
        bool train;
        if (rush) train = pres >= ut_cost(g, UT_WORKER);
        else {
            train = (nworkers < nbases + 1 && pres >= ut_cost(g, UT_WORKER)) || (v2 && nranged > 6);
            int resources = pres;
            if (resources_used != ut_cost(g, UT_BARRACKS) * nbarracks) resources -= ut_cost(g, UT_BARRACKS); // "buffers the resources that are being used for barracks"
            if (building && resources >= ut_cost(g, UT_WORKER) + ut_cost(g, UT_RANGED)) train = true;
        }
        if (train) aa_put(g, i, player, AA_TRAIN, UT_WORKER, 0, 0, REF_NULL, REF_NULL);
    });
    if (pres >= ut_cost(g, UT_RANGED))
        w_for_each(g, n, [&](int i, uint32_t w) { return u_type(w) == UT_BARRACKS && u_pl(w) == pl && a_type(g.a0()[i]) == AT_IDLE; },
                   [&](int i) { aa_put(g, i, player, AA_TRAIN, UT_RANGED, 0, 0, REF_NULL, REF_NULL); });
    w_for_each(g, n, [&](int i, uint32_t w) {
        int fl = ut_flags(g, u_type(w));
        return (fl & UF_ATTACK) && !(fl & UF_HARVEST) && u_pl(w) == pl && a_type(g.a0()[i]) == AT_IDLE;
    }, [&](int i) {
        const bool ranged = u_type(g.w0()[i]) == UT_RANGED;
        if (v2) script_crush2_combat(g, i, player, ranged, rush);
        else if (ranged) script_crush_ranged(g, i, player);
        else script_melee(g, i, player, false);
    });
    return translate_actions(g, player, pathfinder, par0, par1, pn);
}
// EMRDeterministico.getAction (ai/abstraction/EMRDeterministico.java:74-358) followed by translateActions: workers until 4 (6 per
// base once there is a barracks), barracks training Light, Ranged, Heavy in turn, further barracks and bases as resources allow,
// every combat unit attacks the closest enemy.
DEVN int policy_emr(Game &g, int player, int pathfinder, int pn) {
    int par0, par1;
    reserved_resources(g, par0, par1);
    __syncwarp();
    const int n = g.hdr()[H_NUNITS], pl = player + 1, pres = g.hdr()[H_RES0 + player];
    auto count_type = [&](int type) { return w_count(g, n, [&](int, uint32_t w) { return u_pl(w) == pl && u_type(w) == type; }); };
    const int nworkers = count_type(UT_WORKER), nbases = count_type(UT_BASE), nbarracks = count_type(UT_BARRACKS);
    const int narmy = count_type(UT_LIGHT) + count_type(UT_RANGED) + count_type(UT_HEAVY);
    w_for_each(g, n, [&](int i, uint32_t w) { return u_type(w) == UT_BASE && u_pl(w) == pl && a_type(g.a0()[i]) == AT_IDLE; }, [&](int i) { // :133-162
        if (nworkers < (nbarracks == 0 ? 4 : 6 * nbases) && pres >= ut_cost(g, UT_WORKER)) aa_put(g, i, player, AA_TRAIN, UT_WORKER, 0, 0, REF_NULL, REF_NULL);
    });
    w_for_each(g, n, [&](int i, uint32_t w) { return u_type(w) == UT_BARRACKS && u_pl(w) == pl && a_type(g.a0()[i]) == AT_IDLE; }, [&](int i) { // :164-194
        const int turn = narmy % 3, type = turn == 0 ? UT_LIGHT : (turn == 1 ? UT_RANGED : UT_HEAVY);
        if (pres >= ut_cost(g, type)) aa_put(g, i, player, AA_TRAIN, type, 0, 0, REF_NULL, REF_NULL);
    });
    // workers :213-284: every own Worker that can harvest, busy or not, in list order
    auto own_worker = [&](int, uint32_t w) { return u_pl(w) == pl && u_type(w) == UT_WORKER && (ut_flags(g, UT_WORKER) & UF_HARVEST) != 0; };
    const int nw = w_count(g, n, own_worker);
    if (nw > 0) {
        int reserved[4] = {0, 0, 0, 0}, nres = 0, used = 0, taken = 0, wi = -1;
        if (nbases == 0 && taken < nw && pres >= ut_cost(g, UT_BASE) + used) {
            wi = w_next(g, n, wi, own_worker); taken++; script_build_if_not(g, wi, player, UT_BASE, reserved, nres); used += ut_cost(g, UT_BASE);
        }
        if ((nbarracks == 0 || narmy > 2) && taken < nw && pres >= ut_cost(g, UT_BARRACKS) + used) { // the first barracks, or one more
            wi = w_next(g, n, wi, own_worker); taken++; script_build_if_not(g, wi, player, UT_BARRACKS, reserved, nres); used += ut_cost(g, UT_BARRACKS);
        }
        if (nbarracks != 0) {
            // otherResourcePoint :287-311: the resources farther than 10 (in x or in y) from every own base, as a HashSet<Unit>; its
            // first element -- Unit.hashCode() = (int) ID, spread h ^ (h >>> 16) over 16 slots (doubled while size > 0.75 * slots),
            // slots walked upwards, each in insertion order -- is where the next base goes
            auto other = [&](int, uint32_t w) {
                if (!(ut_flags(g, u_type(w)) & UF_RESOURCE)) return false;
                bool mine = false;
                for (int j = 0; j < n; j++) { uint32_t bw = g.w0()[j]; mine |= u_type(bw) == UT_BASE && u_pl(bw) == pl && iabs(u_x(w) - u_x(bw)) <= 10 && iabs(u_y(w) - u_y(bw)) <= 10; }
                return !mine;
            };
            const int no = w_count(g, n, other);
            if (no > 0 && taken < nw && pres >= ut_cost(g, UT_BASE) + used) {
                int cap = 16;
                while (no * 4 > cap * 3) cap *= 2;
                const int first = w_argmin(g, n, [&](int i, uint32_t w) { uint32_t h = g.uid()[i]; h ^= h >> 16; return other(i, w) ? (int)(h & (uint32_t)(cap - 1)) : -1; });
                const uint32_t rw = g.w0()[first];
                wi = w_next(g, n, wi, own_worker); taken++; script_build_if_not(g, wi, player, UT_BASE, reserved, nres, u_x(rw) + 1, u_y(rw) + 1); used += ut_cost(g, UT_BASE);
            }
        }
#pragma unroll 1
        for (int i = w_next(g, n, wi, own_worker); i >= 0; i = w_next(g, n, i, own_worker)) script_harvest(g, i, player, true); // harvestWorkers :325-358
    }
    w_for_each(g, n, [&](int i, uint32_t w) { // :196-211
        int fl = ut_flags(g, u_type(w));
        return (fl & UF_ATTACK) && !(fl & UF_HARVEST) && u_pl(w) == pl && a_type(g.a0()[i]) == AT_IDLE;
    }, [&](int i) { script_melee(g, i, player, false); });
    return translate_actions(g, player, pathfinder, par0, par1, pn);
}
#endif

// WorkerRush.getAction / LightRush.getAction followed by translateActions; appends to the pending list from pn.
// Lane 0 does the work; returns the new pending count (uniform).
DEVN int policy_scripted(Game &g, int player, int kind, int pathfinder, int pn) {
    // the PO rushes are their rush plus exploration, which only a partially observable view triggers (`gs instanceof
    // PartiallyObservableGameState`, POLightRush.java:57)
#ifdef MRTS_TU_RUSH_ONLY
    const bool explore = false;
#else
    const bool explore = POL_IS_PO_RUSH(kind) && g.po_view;
    if (POL_IS_PO_RUSH(kind)) kind = kind - POL_PO_WORKER_RUSH + POL_WORKER_RUSH;
#endif
#ifndef MRTS_TU_RUSH_ONLY
    if (kind == POL_EMR_DETERMINISTICO) return policy_emr(g, player, pathfinder, pn);
    if (kind == POL_CRUSH_V1 || kind == POL_CRUSH_V2) return policy_crush(g, player, pathfinder, pn, kind == POL_CRUSH_V2);
#endif
    int par0, par1;
    reserved_resources(g, par0, par1);
    __syncwarp();
    const int n = g.hdr()[H_NUNITS], pl = player + 1, pres = g.hdr()[H_RES0 + player];
    // the barracks scripts: LightRush / LightDefense, and Heavy* / Ranged* = the same classes with the trained type swapped;
    // the defenses share their rush's skeleton and differ in script_melee / script_harvest
#ifdef MRTS_TU_RUSH_ONLY
    const bool light = kind != POL_WORKER_RUSH, defense = false, always = false;
#else
    const bool light = kind != POL_WORKER_RUSH && kind != POL_WORKER_DEFENSE && kind != POL_WORKER_RUSH_PP, defense = POL_IS_DEFENSE(kind), always = kind == POL_WORKER_RUSH_PP;
#endif
    const int UT_RUSH = (kind == POL_HEAVY_RUSH || kind == POL_HEAVY_DEFENSE) ? 5 : ((kind == POL_RANGED_RUSH || kind == POL_RANGED_DEFENSE) ? 6 : UT_LIGHT); // HeavyRush.java:55, RangedRush.java:52
    auto own_harvester = [&](int, uint32_t w) { return u_pl(w) == pl && (ut_flags(g, u_type(w)) & UF_HARVEST) != 0; };
    // bases (WorkerRush.java:70-76,100-102; LightRush.java:83-89,123-133)
    w_for_each(g, n, [&](int i, uint32_t w) { return u_type(w) == UT_BASE && u_pl(w) == pl && a_type(g.a0()[i]) == AT_IDLE; }, [&](int i) {
        bool train = pres >= ut_cost(g, UT_WORKER);
        if (light && train) train = w_count(g, n, [&](int, uint32_t ow) { return u_type(ow) == UT_WORKER && u_pl(ow) == pl; }) < 1;
        if (train) aa_put(g, i, player, AA_TRAIN, UT_WORKER, 0, 0, REF_NULL, REF_NULL);
    });
    if (light) { // barracks (LightRush.java:92-98,135-139)
        if (pres >= ut_cost(g, UT_RUSH))
            w_for_each(g, n, [&](int i, uint32_t w) { return u_type(w) == UT_BARRACKS && u_pl(w) == pl && a_type(g.a0()[i]) == AT_IDLE; },
                       [&](int i) { aa_put(g, i, player, AA_TRAIN, UT_RUSH, 0, 0, REF_NULL, REF_NULL); });
    }
    w_for_each(g, n, [&](int i, uint32_t w) { // melee units
        int fl = ut_flags(g, u_type(w));
        return (fl & UF_ATTACK) && !(fl & UF_HARVEST) && u_pl(w) == pl && a_type(g.a0()[i]) == AT_IDLE;
    }, [&](int i) { script_melee(g, i, player, defense, explore, always); });
    // workers: all own harvesters, busy ones too, in list order
    int nbases = w_count(g, n, [&](int, uint32_t w) { return u_pl(w) == pl && u_type(w) == UT_BASE; });
    int nbarracks = w_count(g, n, [&](int, uint32_t w) { return u_pl(w) == pl && u_type(w) == UT_BARRACKS; });
    int nworkers = w_count(g, n, own_harvester);
    if (nworkers > 0) {
        int reserved[4] = {0, 0, 0, 0}, nres = 0, used = 0, taken = 0; // `taken` workers were removed from the front of freeWorkers
        int wi = -1;                                     // cursor over own harvesters in list order
        if (nbases == 0 && taken < nworkers) {
            if (pres >= ut_cost(g, UT_BASE) + used) { wi = w_next(g, n, wi, own_harvester); taken++; script_build_if_not(g, wi, player, UT_BASE, reserved, nres); used += ut_cost(g, UT_BASE); }
        }
        if (light) {
            if (nbarracks == 0 && pres >= ut_cost(g, UT_BARRACKS) + used && taken < nworkers) {
                wi = w_next(g, n, wi, own_harvester); taken++; script_build_if_not(g, wi, player, UT_BARRACKS, reserved, nres); used += ut_cost(g, UT_BARRACKS);
            }
            // harvest with every remaining worker; those that cannot, attack -- in a second pass, as the reference does
            uint32_t still[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll 1
            for (int base = 0; base < n; base += 32) { // (one ballot per 32 units instead of a search per worker)
                int i = base + g.lane;
                unsigned mask = __ballot_sync(FULLM, i < n && i > wi && own_harvester(i, g.w0()[i])), st = 0;
#pragma unroll 1
                while (mask) { int b = __ffs(mask) - 1; mask &= mask - 1; if (script_harvest(g, base + b, player, defense)) st |= 1u << b; }
                still[base >> 5] = st;
            }
#pragma unroll 1
            for (int base = 0; base < n; base += 32) {
                unsigned mask = still[base >> 5];
#pragma unroll 1
                while (mask) { int b = __ffs(mask) - 1; mask &= mask - 1; script_melee(g, base + b, player, defense, explore, always); }
            }
        } else {
            // WorkerRush.java:146-202: one harvester, the rest attack; a harvester that stays free is appended at the END
            int hw = -1;
            if (taken < nworkers) { hw = w_next(g, n, wi, own_harvester); wi = hw; taken++; }
            bool hw_free = hw >= 0 && script_harvest(g, hw, player, defense);
#pragma unroll 1
            for (int base = 0; base < n; base += 32) {
                int i = base + g.lane;
                unsigned mask = __ballot_sync(FULLM, i < n && i > wi && own_harvester(i, g.w0()[i]));
#pragma unroll 1
                while (mask) { int b = __ffs(mask) - 1; mask &= mask - 1; script_melee(g, base + b, player, defense, explore, always); }
            }
            if (hw_free) script_melee(g, hw, player, defense, explore, always);
        }
    }
    return translate_actions(g, player, pathfinder, par0, par1, pn);
}
