package rts.cuda;

import java.lang.foreign.Arena;
import java.lang.foreign.FunctionDescriptor;
import java.lang.foreign.Linker;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.SymbolLookup;
import java.lang.foreign.ValueLayout;
import java.lang.invoke.MethodHandle;
import java.nio.file.Path;

import rts.units.UnitTypeTable;

/**
 * n independent microRTS games stepped in lockstep on one GPU through libmicrorts_cuda.so
 * (include/microrts_cuda.h).  Mirrors the part of rts.GameState / tests.JNIGridnetVecClient that the batched
 * simulation path needs: reset, issueSafe/issue, cycle, gameStep, getVectorObservation, getMasks, winner/gameover.
 *
 * Binding: Panama FFM (java.lang.foreign, JDK 22+).  NOT COMPILED in the build container of this repository (no JDK
 * there); the signatures are written against include/microrts_cuda.h.  A JNI variant needs only a 1:1 C shim.
 *
 * Threading: like a GameState, an instance is confined to one thread.  One instance per GPU.
 */
public final class BatchedGameState implements AutoCloseable {
    public static final int POLICY_EXTERNAL = 0, POLICY_PASSIVE = 1, POLICY_RANDOM_BIASED = 2, POLICY_WORKER_RUSH = 3, POLICY_LIGHT_RUSH = 4,
            POLICY_HEAVY_RUSH = 5, POLICY_RANGED_RUSH = 6, POLICY_WORKER_DEFENSE = 7, POLICY_LIGHT_DEFENSE = 8, POLICY_HEAVY_DEFENSE = 9,
            POLICY_RANGED_DEFENSE = 10, POLICY_PO_WORKER_RUSH = 11, POLICY_PO_LIGHT_RUSH = 12, POLICY_PO_HEAVY_RUSH = 13,
            POLICY_PO_RANGED_RUSH = 14, POLICY_WORKER_RUSH_PP = 15, POLICY_CRUSH_V1 = 16, POLICY_CRUSH_V2 = 17, POLICY_EMR_DETERMINISTICO = 18;   // MRTS_POLICY_* of include/microrts_cuda.h
    public static final int PF_ASTAR = 0, PF_BFS = 1, PF_GREEDY = 2, PF_FLOODFILL = 3; // ai.abstraction.pathfinding.{AStar,BFS,Greedy,FloodFill}PathFinding
    public static final int FLAG_PARTIAL_OBS = 1, FLAG_SCRIPTED_AI = 2, FLAG_PO_POLICIES = 4; // MRTS_FLAG_*
    public static final int ACTIONS_VECTOR = 0, ACTIONS_RAW = 1;
    public static final int DTYPE_U8 = 0, DTYPE_I32 = 1, DTYPE_BITS = 2;
    public static final int NCCL_UNIQUE_ID_BYTES = 128, PAG_DONE = -100;

    private static final Linker LINKER = Linker.nativeLinker();
    private static final ValueLayout.OfInt I = ValueLayout.JAVA_INT;
    private static final ValueLayout.OfLong L = ValueLayout.JAVA_LONG;
    private static final java.lang.foreign.AddressLayout P = ValueLayout.ADDRESS;

    private final Arena arena = Arena.ofConfined();
    private final SymbolLookup lib;
    private final MethodHandle lastError, uttCreate, uttDestroy, mapLoad, mapDestroy, batchCreate, batchDestroy, reset, resetMasked,
            setPolicy, setAutoReset, setActions, issue, step, cycleTo, observe, masks, results, stats, rollout, pathfind, evaluate, numPlanes, maskWidth,
            restartMasked, setIssueOrder, setInfoOutput, setObservationOutputs, setMaskOutputs, setOutputStride, setVecAutoreset, setActionsInterleaved,
            vecStep, copyGames, scatterGames, copyToHost, sync, unitActions, cycleToDecision, playerActions, statsAllreduce, ncclUniqueId, ncclCommCreate, ncclCommDestroy,
            mctsCreate, mctsIterate, mctsRoot, mctsBestActions, mctsDestroy, pagCreate, pagNext, pagSize, pagDestroy;

    private MemorySegment utt, map, batch;
    public final int numGames, width, height, planes, maskW;

    private MethodHandle h(String name, FunctionDescriptor d) {
        return LINKER.downcallHandle(lib.find(name).orElseThrow(() -> new UnsatisfiedLinkError(name)), d);
    }

    /**
     * @param libraryPath path of libmicrorts_cuda.so
     * @param uttVersion  UnitTypeTable.VERSION_* ; @param conflictPolicy UnitTypeTable.MOVE_CONFLICT_RESOLUTION_*
     * @param mapPath     a map XML as accepted by PhysicalGameState.load
     */
    public BatchedGameState(Path libraryPath, int uttVersion, int conflictPolicy, String mapPath, int numGames, int device,
                            boolean partialObs) {
        this(libraryPath, uttVersion, conflictPolicy, mapPath, numGames, device, partialObs ? FLAG_PARTIAL_OBS : 0);
    }

    /**
     * @param flags FLAG_PARTIAL_OBS (8-plane observations), FLAG_SCRIPTED_AI (the scripted POLICY_* and findPath need it),
     *              FLAG_PO_POLICIES (Game with partiallyObservable = true: device policies decide on their player's view)
     */
    public BatchedGameState(Path libraryPath, int uttVersion, int conflictPolicy, String mapPath, int numGames, int device,
                            int flags) {
        lib = SymbolLookup.libraryLookup(libraryPath, arena);
        lastError = h("mrts_last_error", FunctionDescriptor.of(P));
        uttCreate = h("mrts_utt_create", FunctionDescriptor.of(I, I, I, P));
        uttDestroy = h("mrts_utt_destroy", FunctionDescriptor.ofVoid(P));
        mapLoad = h("mrts_map_load_xml", FunctionDescriptor.of(I, P, P, P));
        mapDestroy = h("mrts_map_destroy", FunctionDescriptor.ofVoid(P));
        batchCreate = h("mrts_batch_create", FunctionDescriptor.of(I, P, P, I, L, I, I, I, P));
        batchDestroy = h("mrts_batch_destroy", FunctionDescriptor.ofVoid(P));
        reset = h("mrts_batch_reset", FunctionDescriptor.of(I, P, P, I));
        resetMasked = h("mrts_batch_reset_masked", FunctionDescriptor.of(I, P, P, P, I));
        setPolicy = h("mrts_batch_set_policy", FunctionDescriptor.of(I, P, I, I, I));
        setAutoReset = h("mrts_batch_set_auto_reset", FunctionDescriptor.of(I, P, I));
        setActions = h("mrts_batch_set_actions", FunctionDescriptor.of(I, P, I, I, P, P, I, I, I));
        issue = h("mrts_batch_issue", FunctionDescriptor.of(I, P, I, I, P, P, I, I, I, I));
        step = h("mrts_batch_step", FunctionDescriptor.of(I, P, I, I));
        cycleTo = h("mrts_batch_cycle_to", FunctionDescriptor.of(I, P, P, I, I));
        observe = h("mrts_batch_observe", FunctionDescriptor.of(I, P, I, I, P, I));
        masks = h("mrts_batch_masks", FunctionDescriptor.of(I, P, I, I, P, I));
        results = h("mrts_batch_results", FunctionDescriptor.of(I, P, P, I));
        stats = h("mrts_batch_stats", FunctionDescriptor.of(I, P, P));
        rollout = h("mrts_batch_rollout", FunctionDescriptor.of(I, P, I, I, I, I, I, P, P, P, I));
        pathfind = h("mrts_batch_pathfind", FunctionDescriptor.of(I, P, I, P, P, I));
        evaluate = h("mrts_batch_evaluate", FunctionDescriptor.of(I, P, I, I, I, P, I));
        numPlanes = h("mrts_batch_num_planes", FunctionDescriptor.of(I, P));
        maskWidth = h("mrts_batch_mask_width", FunctionDescriptor.of(I, P));
        restartMasked = h("mrts_batch_restart_masked", FunctionDescriptor.of(I, P, P, I));
        setIssueOrder = h("mrts_batch_set_issue_order", FunctionDescriptor.of(I, P, I));
        setInfoOutput = h("mrts_batch_set_info_output", FunctionDescriptor.of(I, P, P));
        setObservationOutputs = h("mrts_batch_set_observation_outputs", FunctionDescriptor.of(I, P, I, P, P));
        setMaskOutputs = h("mrts_batch_set_mask_outputs", FunctionDescriptor.of(I, P, P, P));
        setOutputStride = h("mrts_batch_set_output_stride", FunctionDescriptor.of(I, P, I));
        setVecAutoreset = h("mrts_batch_set_vec_autoreset", FunctionDescriptor.of(I, P, I, I));
        setActionsInterleaved = h("mrts_batch_set_actions_interleaved", FunctionDescriptor.of(I, P, I, P, I, I, I, I));
        vecStep = h("mrts_batch_vec_step", FunctionDescriptor.of(I, P, P, I, I, I));
        copyGames = h("mrts_batch_copy_games", FunctionDescriptor.of(I, P, P, P, P, I));
        scatterGames = h("mrts_batch_scatter_games", FunctionDescriptor.of(I, P, P, P, I));
        copyToHost = h("mrts_batch_copy_to_host", FunctionDescriptor.of(I, P, P, P, L));
        sync = h("mrts_batch_sync", FunctionDescriptor.of(I, P));
        unitActions = h("mrts_batch_unit_actions", FunctionDescriptor.of(I, P, I, I, I, I, P, P, P, P, I));
        cycleToDecision = h("mrts_batch_cycle_to_decision", FunctionDescriptor.of(I, P));
        playerActions = h("mrts_batch_player_actions", FunctionDescriptor.of(I, P, L, I, P, P, L, I, P));
        statsAllreduce = h("mrts_batch_stats_allreduce", FunctionDescriptor.of(I, P, P, P));
        ncclUniqueId = h("mrts_nccl_unique_id", FunctionDescriptor.of(I, P));
        ncclCommCreate = h("mrts_nccl_comm_create", FunctionDescriptor.of(I, P, I, I, I, P));
        ncclCommDestroy = h("mrts_nccl_comm_destroy", FunctionDescriptor.ofVoid(P));
        mctsCreate = h("mrts_mcts_create", FunctionDescriptor.of(I, P, I, P, I, P, P));
        mctsIterate = h("mrts_mcts_iterate", FunctionDescriptor.of(I, P, I));
        mctsRoot = h("mrts_mcts_root", FunctionDescriptor.of(I, P, L, P, P, P, P, I));
        mctsBestActions = h("mrts_mcts_best_actions", FunctionDescriptor.of(I, P, P, P, I));
        mctsDestroy = h("mrts_mcts_destroy", FunctionDescriptor.ofVoid(P));
        pagCreate = h("mrts_pag_create", FunctionDescriptor.of(I, P, L, I, I, P));
        pagNext = h("mrts_pag_next", FunctionDescriptor.of(I, P, P, I));
        pagSize = h("mrts_pag_size", FunctionDescriptor.of(L, P));
        pagDestroy = h("mrts_pag_destroy", FunctionDescriptor.ofVoid(P));
        try {
            MemorySegment out = arena.allocate(P);
            check((int) uttCreate.invoke(uttVersion, conflictPolicy, out));
            utt = out.get(P, 0);
            check((int) mapLoad.invoke(arena.allocateFrom(mapPath), utt, out));
            map = out.get(P, 0);
            MemorySegment maps = arena.allocate(P);
            maps.set(P, 0, map);
            check((int) batchCreate.invoke(utt, maps, 1, (long) numGames, device, flags, 0, out));
            batch = out.get(P, 0);
            this.numGames = numGames;
            this.planes = (int) numPlanes.invoke(batch);
            this.maskW = (int) maskWidth.invoke(batch);
            // width/height come from the map: mrts_map_width / mrts_map_height (omitted handles for brevity)
            this.width = (int) h("mrts_map_width", FunctionDescriptor.of(I, P)).invoke(map);
            this.height = (int) h("mrts_map_height", FunctionDescriptor.of(I, P)).invoke(map);
        } catch (Throwable t) {
            throw new RuntimeException(t);
        }
    }

    private void check(int rc) throws Throwable {
        if (rc < 0) {
            MemorySegment msg = ((MemorySegment) lastError.invoke()).reinterpret(4096);
            throw new IllegalStateException("libmicrorts_cuda error " + rc + ": " + msg.getString(0));
        }
    }

    /** new GameState(pgs, utt) for every game; seeds[g] seeds game g's java.util.Random streams (null: seed = g). */
    public void reset(long[] seeds) throws Throwable {
        try (Arena a = Arena.ofConfined()) {
            check((int) reset.invoke(batch, seeds == null ? MemorySegment.NULL : a.allocateFrom(L, seeds), 0));
        }
    }

    /** Which AI's getAction runs inside the step kernel for `player` (POLICY_*). */
    public void setPolicy(int player, int policy) throws Throwable { check((int) setPolicy.invoke(batch, player, policy, 0)); }

    /** JNIGridnetVecClient auto-reset (src/tests/JNIGridnetVecClient.java:272-286), done on the device. */
    public void setAutoReset(boolean on) throws Throwable { check((int) setAutoReset.invoke(batch, on ? 1 : 0)); }

    /** Restart the games whose mask byte is set; their RNG streams keep running (JNIGridnetVecClient.java:272-286). */
    public void restartMasked(byte[] mask) throws Throwable {
        try (Arena a = Arena.ofConfined()) {
            check((int) restartMasked.invoke(batch, a.allocateFrom(ValueLayout.JAVA_BYTE, mask), 0));
        }
    }

    /** true: player 1's PlayerAction is built after player 0's is issued (JNIGridnetClientSelfPlay.java:160-170). */
    public void setSequentialIssue(boolean on) throws Throwable { check((int) setIssueOrder.invoke(batch, on ? 1 : 0)); }

    /**
     * Device buffers (CUDA device pointers wrapped as MemorySegment.ofAddress) that every later step() fills: the reward
     * functions' step facts [numGames][2][12] and the observations of player 0 / 1 [numGames][planes][height][width].
     */
    public void setDeviceOutputs(MemorySegment info, int dtype, MemorySegment obs0, MemorySegment obs1) throws Throwable {
        check((int) setInfoOutput.invoke(batch, info));
        check((int) setObservationOutputs.invoke(batch, dtype, obs0, obs1));
    }

    /** Stage vector actions [numGames][maxK][8] (PlayerAction.fromVectorAction rows) for an EXTERNAL player. */
    public void setVectorActions(int player, int[] actions, int[] counts, int maxK) throws Throwable {
        try (Arena a = Arena.ofConfined()) {
            check((int) setActions.invoke(batch, player, ACTIONS_VECTOR, a.allocateFrom(I, actions), a.allocateFrom(I, counts), maxK, 1, 0));
        }
    }

    /** gs.issueSafe(pa) for every game: rows [cell, type, parameter, x, y, unitType, 0, 0]. */
    public void issueSafe(int player, int[] rawActions, int[] counts, int maxK) throws Throwable {
        try (Arena a = Arena.ofConfined()) {
            check((int) issue.invoke(batch, player, ACTIONS_RAW, a.allocateFrom(I, rawActions), a.allocateFrom(I, counts), maxK, -1, 1, 0));
        }
    }

    /** Game.start loop body (src/rts/Game.java:126-140) nCycles times: policies, issueSafe x2, cycle. */
    public void step(int nCycles, int maxCycles) throws Throwable { check((int) step.invoke(batch, nCycles, maxCycles)); }

    /** gs.cycle() nCycles times, no policies. */
    public void cycle(int nCycles) throws Throwable { check((int) cycleTo.invoke(batch, MemorySegment.NULL, nCycles, 0)); }

    /** gs.getVectorObservation(player) of every game into out[numGames][planes][height][width]. */
    public void getVectorObservation(int player, int[] out) throws Throwable {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment seg = a.allocate(I, out.length);
            check((int) observe.invoke(batch, player, DTYPE_I32, seg, 0));
            MemorySegment.copy(seg, I, 0, out, 0, out.length);
        }
    }

    /** JNIGridnetClient.getMasks(player) of every game into out[numGames][height][width][maskW]. */
    public void getMasks(int player, int[] out) throws Throwable {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment seg = a.allocate(I, out.length);
            check((int) masks.invoke(batch, player, DTYPE_I32, seg, 0));
            MemorySegment.copy(seg, I, 0, out, 0, out.length);
        }
    }

    /** out[g] = {gs.getTime(), gs.winner(), gs.gameover() ? 1 : 0, error bits}. */
    public int[] results() throws Throwable {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment seg = a.allocate(I, 4L * numGames);
            check((int) results.invoke(batch, seg, 0));
            return seg.toArray(I);
        }
    }

    /** {wins_p0, wins_p1, draws, games_finished, cycles, decisions, unit_cycles, errors} since the last reset. */
    public long[] stats() throws Throwable {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment seg = a.allocate(L, 8);
            check((int) stats.invoke(batch, seg));
            return seg.toArray(L);
        }
    }

    /**
     * NaiveMCTS.simulate + ef.evaluate for rolloutsPerGame clones of every game (src/ai/mcts/naivemcts/NaiveMCTS.java:195-223).
     * Returns the undiscounted evaluations; multiply by Math.pow(0.99, time/10.0) with outTime as the reference does.
     */
    public float[] rollout(int rolloutsPerGame, int depth, int evalFn, int maxplayer, int observer, long[] seeds, int[] outTime) throws Throwable {
        try (Arena a = Arena.ofConfined()) {
            long n = (long) numGames * rolloutsPerGame;
            MemorySegment ev = a.allocate(ValueLayout.JAVA_FLOAT, n), tm = a.allocate(I, n);
            check((int) rollout.invoke(batch, rolloutsPerGame, depth, evalFn, maxplayer, observer,
                    seeds == null ? MemorySegment.NULL : a.allocateFrom(L, seeds), ev, tm, 0));
            if (outTime != null) MemorySegment.copy(tm, I, 0, outTime, 0, outTime.length);
            return ev.toArray(ValueLayout.JAVA_FLOAT);
        }
    }

    /** EvaluationFunction.evaluate(maxplayer, 1 - maxplayer, gs) of every game (evalFn 0 = SimpleSqrtEvaluationFunction3, 1 = SimpleEvaluationFunction). */
    public float[] evaluate(int evalFn, int maxplayer, int observer) throws Throwable {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment out = a.allocate(ValueLayout.JAVA_FLOAT, numGames);
            check((int) evaluate.invoke(batch, evalFn, maxplayer, observer, out, 0));
            return out.toArray(ValueLayout.JAVA_FLOAT);
        }
    }

    /**
     * PathFinding.findPathToPositionInRange(start, targetpos, range, gs, null) for one unit per game
     * (src/ai/abstraction/pathfinding/PathFinding.java:17-24): queries = {cell of the unit, target position, range} per game,
     * pathfinder = PF_ASTAR / PF_BFS / PF_GREEDY.  Returns the direction of the MOVE per game, -1 for null.
     */
    public int[] findPath(int pathfinder, int[] queries) throws Throwable {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment out = a.allocate(I, numGames);
            check((int) pathfind.invoke(batch, pathfinder, a.allocateFrom(I, queries), out, 0));
            return out.toArray(I);
        }
    }

    // ---- the vectorised RL flow in one launch per gameStep (tests.CudaGridnetVecClient drives these) ---------------------------------
    /** Fused outputs in JNIGridnetVecClient's environment order: obs = [2 * numGames][planes][h][w] device memory, environment 2g = player 0 of game g. */
    public void setInterleavedOutputs(int dtype, MemorySegment obs, long obsBytesPerEnv, MemorySegment masks, long maskBytesPerEnv) throws Throwable {
        check((int) setOutputStride.invoke(batch, 2));
        check((int) setObservationOutputs.invoke(batch, dtype, obs, obs.asSlice(obsBytesPerEnv)));
        check((int) setMaskOutputs.invoke(batch, masks, masks.equals(MemorySegment.NULL) ? MemorySegment.NULL : masks.asSlice(maskBytesPerEnv)));
    }

    /** JNIGridnetVecClient's auto-reset inside the step launch: doneMode 1 game over, 2 resources exhausted, 3 never; always after maxSteps. */
    public void setVecAutoreset(int doneMode, int maxSteps) throws Throwable { check((int) setVecAutoreset.invoke(batch, doneMode, maxSteps)); }

    /** Both players' vector actions of every game from one [2 * numGames][maxK][8] array in environment order (off-heap, ideally pinned). */
    public void setActionsInterleaved(MemorySegment actions, int maxK, boolean async) throws Throwable {
        check((int) setActionsInterleaved.invoke(batch, ACTIONS_VECTOR, actions, maxK, 1, 0, async ? 1 : 0));
    }

    /** setActionsInterleaved + step(1) in one native call: JNIGridnetVecClient.gameStep of the self-play environments. */
    public void vecStep(MemorySegment actions, int maxK, boolean async) throws Throwable { check((int) vecStep.invoke(batch, actions, maxK, 0, async ? 1 : 0)); }

    /** Queue a device -> host copy behind the step on the batch's stream; sync() waits. */
    public void copyToHost(MemorySegment host, MemorySegment device, long bytes) throws Throwable { check((int) copyToHost.invoke(batch, host, device, bytes)); }

    public void sync() throws Throwable { check((int) sync.invoke(batch)); }

    /** gs.clone() for many games at once: game g becomes a copy of src's game srcIndex[g] (null: g). */
    public void copyGames(BatchedGameState src, long[] srcIndex) throws Throwable {
        try (Arena a = Arena.ofConfined()) {
            check((int) copyGames.invoke(batch, src.batch, srcIndex == null ? MemorySegment.NULL : a.allocateFrom(L, srcIndex), MemorySegment.NULL, 0));
        }
    }

    /** cycle() until the game is over or somebody can act (the loop at the head of NaiveMCTSNode). */
    public void cycleToDecision() throws Throwable { check((int) cycleToDecision.invoke(batch)); }

    /** GameState.getPlayerActions(player) of one game as RAW rows; returns the number of PlayerActions. */
    public long getPlayerActions(long game, int player, int[] outRows, int[] outCounts, int maxK) throws Throwable {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment rows = a.allocate(I, outRows.length), counts = a.allocate(I, outCounts.length), total = a.allocate(L);
            check((int) playerActions.invoke(batch, game, player, rows, counts, (long) outCounts.length, maxK, total));
            MemorySegment.copy(rows, I, 0, outRows, 0, outRows.length);
            MemorySegment.copy(counts, I, 0, outCounts, 0, outCounts.length);
            return total.get(L, 0);
        }
    }

    /**
     * One NaiveMCTS search per game (src/ai/mcts/naivemcts/NaiveMCTS.java): `iterations` playouts each, then getBestActionSoFar as RAW rows
     * [numGames][maxK][8] + counts -- issue them with issueSafe(player, rows, counts, maxK).
     */
    public int[] naiveMcts(int player, int iterations, int lookahead, int maxDepth, float eL, float eG, float e0, long[] seeds, int[] outCounts, int maxK) throws Throwable {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment prm = a.allocate(32);
            prm.set(I, 0, lookahead); prm.set(I, 4, maxDepth); prm.set(ValueLayout.JAVA_FLOAT, 8, eL); prm.set(ValueLayout.JAVA_FLOAT, 12, eG);
            prm.set(ValueLayout.JAVA_FLOAT, 16, e0); prm.set(I, 20, 0); prm.set(I, 24, 1); prm.set(I, 28, 0);
            MemorySegment out = a.allocate(P);
            check((int) mctsCreate.invoke(batch, player, prm, iterations + 2, seeds == null ? MemorySegment.NULL : a.allocateFrom(L, seeds), out));
            MemorySegment m = out.get(P, 0);
            try {
                check((int) mctsIterate.invoke(m, iterations));
                MemorySegment rows = a.allocate(I, (long) numGames * maxK * 8), counts = a.allocate(I, numGames);
                check((int) mctsBestActions.invoke(m, rows, counts, maxK));
                MemorySegment.copy(counts, I, 0, outCounts, 0, numGames);
                return rows.toArray(I);
            } finally {
                mctsDestroy.invoke(m);
            }
        }
    }

    /** The run's one collective: the eight counters summed over every rank's batch (ncclAllReduce inside the library). */
    public long[] statsAllReduce(MemorySegment comm) throws Throwable {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment seg = a.allocate(L, 8);
            check((int) statsAllreduce.invoke(batch, seg, comm));
            return seg.toArray(L);
        }
    }

    /** Rank 0 of a multi-GPU host calls this and hands the 128 bytes to the other ranks (threads: an array; processes: a file or socket). */
    public byte[] ncclUniqueId() throws Throwable {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment id = a.allocate(NCCL_UNIQUE_ID_BYTES);
            check((int) ncclUniqueId.invoke(id));
            return id.toArray(ValueLayout.JAVA_BYTE);
        }
    }

    /** Collective over all ranks: the communicator statsAllReduce takes. */
    public MemorySegment ncclCommCreate(byte[] uniqueId, int nRanks, int rank, int device) throws Throwable {
        MemorySegment out = arena.allocate(P);
        check((int) ncclCommCreate.invoke(arena.allocateFrom(ValueLayout.JAVA_BYTE, uniqueId), nRanks, rank, device, out));
        return out.get(P, 0);
    }

    @Override
    public void close() {
        try {
            if (batch != null) batchDestroy.invoke(batch);
            if (map != null) mapDestroy.invoke(map);
            if (utt != null) uttDestroy.invoke(utt);
        } catch (Throwable ignored) {
        }
        arena.close();
    }

    /** Convenience: does `utt` (a reference UnitTypeTable) match the built-in table the native side was created with? */
    public static boolean sameTable(UnitTypeTable utt, int version) {
        UnitTypeTable ref = new UnitTypeTable(version, utt.getMoveConflictResolutionStrategy());
        return ref.getUnitTypes().size() == utt.getUnitTypes().size();
    }
}
