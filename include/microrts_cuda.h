/*
 * microrts_cuda.h -- C ABI of libmicrorts_cuda.so, the B200-native batched microRTS simulator.
 *
 * The reference (ConnAALL/MicroRTS) has no FFI: its seam is the Java object API.  Each entry point below names
 * the reference interface it replaces (file:line under the reference checkout).  A new Java class
 * rts.cuda.BatchedGameState binds these symbols through Panama FFM or JNI; the Python
 * package microrts_b200 binds them through ctypes.
 *
 * Conventions
 *  - Every function returns 0 (MRTS_OK) or a negative MRTS_E_* code.  Nothing throws, nothing prints.
 *    mrts_last_error() returns a human readable description of the last failure on the calling thread.
 *  - Handles are opaque and library-owned; array arguments are caller-owned and never retained after return.
 *  - A mrts_batch is bound to one CUDA device and one stream and is NOT thread-safe (one host thread per batch,
 *    one batch per GPU; games shard across GPUs by contiguous index range, no data-path collective).
 *  - "on_device" pointers are CUDA device pointers on the batch's device; otherwise host pointers.
 *  - There is no CPU fallback: every compute entry point fails with MRTS_E_CUDA when no device is usable.
 *  - Per-game anomalies (the reference's `throw new Error`, System.err chatter) set sticky bits in the game's
 *    error word, readable through mrts_batch_export / mrts_batch_results.
 */
#ifndef MICRORTS_CUDA_H
#define MICRORTS_CUDA_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MRTS_ABI_VERSION 2

enum {
    MRTS_OK = 0,
    MRTS_E_ARG = -1,      /* bad argument */
    MRTS_E_IO = -2,       /* file not found / unreadable */
    MRTS_E_PARSE = -3,    /* malformed XML / JSON */
    MRTS_E_CUDA = -4,     /* CUDA runtime error (no device, launch failure, out of memory) */
    MRTS_E_LIMIT = -5,    /* map or unit table exceeds an engine limit */
    MRTS_E_STATE = -6     /* call not valid in the current state */
};

/* UnitAction.TYPE_*  (src/rts/UnitAction.java:29-59) */
enum { MRTS_NONE = 0, MRTS_MOVE = 1, MRTS_HARVEST = 2, MRTS_RETURN = 3, MRTS_PRODUCE = 4, MRTS_ATTACK = 5 };
/* UnitAction.DIRECTION_*  (src/rts/UnitAction.java:68-100) */
enum { MRTS_DIR_NONE = -1, MRTS_UP = 0, MRTS_RIGHT = 1, MRTS_DOWN = 2, MRTS_LEFT = 3 };
/* UnitTypeTable.VERSION_* / MOVE_CONFLICT_RESOLUTION_*  (src/rts/units/UnitTypeTable.java:28-60) */
enum { MRTS_UTT_ORIGINAL = 1, MRTS_UTT_FINETUNED = 2, MRTS_UTT_NON_DETERMINISTIC = 3 };
enum { MRTS_CANCEL_BOTH = 1, MRTS_CANCEL_RANDOM = 2, MRTS_CANCEL_ALTERNATING = 3 };

/* Device-resident policies (the AI classes whose getAction() runs inside the step kernel). */
enum {
    MRTS_POLICY_EXTERNAL = 0,      /* actions staged with mrts_batch_set_actions (JNIAI, src/ai/jni/JNIAI.java:51-55) */
    MRTS_POLICY_PASSIVE = 1,       /* ai.PassiveAI: never issues anything */
    MRTS_POLICY_RANDOM_BIASED = 2, /* ai.RandomBiasedAI (src/ai/RandomBiasedAI.java:51-107) */
    MRTS_POLICY_WORKER_RUSH = 3,   /* ai.abstraction.WorkerRush (src/ai/abstraction/WorkerRush.java:63-204) */
    MRTS_POLICY_LIGHT_RUSH = 4,    /* ai.abstraction.LightRush  (src/ai/abstraction/LightRush.java:77-258) */
    MRTS_POLICY_HEAVY_RUSH = 5,    /* ai.abstraction.HeavyRush  (src/ai/abstraction/HeavyRush.java: LightRush training Heavy units) */
    MRTS_POLICY_RANGED_RUSH = 6,   /* ai.abstraction.RangedRush (src/ai/abstraction/RangedRush.java: LightRush training Ranged units) */
    MRTS_POLICY_WORKER_DEFENSE = 7, /* ai.abstraction.WorkerDefense (src/ai/abstraction/WorkerDefense.java:74-209) */
    MRTS_POLICY_LIGHT_DEFENSE = 8,  /* ai.abstraction.LightDefense  (src/ai/abstraction/LightDefense.java:78-247) */
    MRTS_POLICY_HEAVY_DEFENSE = 9,  /* ai.abstraction.HeavyDefense  (src/ai/abstraction/HeavyDefense.java: LightDefense training Heavy units) */
    MRTS_POLICY_RANGED_DEFENSE = 10, /* ai.abstraction.RangedDefense (src/ai/abstraction/RangedDefense.java: LightDefense training Ranged units) */
    /* src/ai/abstraction/partialobservability/PO{Worker,Light,Heavy,Ranged}Rush.java: the rush, whose idle combat units walk to
       the nearest cell outside their player's sight when no enemy is visible (MRTS_FLAG_PO_POLICIES batches; without the flag
       they behave as the plain rushes, as the reference classes do on a fully observable GameState) */
    MRTS_POLICY_PO_WORKER_RUSH = 11, MRTS_POLICY_PO_LIGHT_RUSH = 12, MRTS_POLICY_PO_HEAVY_RUSH = 13, MRTS_POLICY_PO_RANGED_RUSH = 14,
    MRTS_POLICY_WORKER_RUSH_PP = 15, /* ai.abstraction.WorkerRushPlusPlus (src/ai/abstraction/WorkerRushPlusPlus.java:74-207) */
    MRTS_POLICY_CRUSH_V1 = 16, /* ai.abstraction.cRush.CRush_V1 (src/ai/abstraction/cRush/CRush_V1.java:68-421, RangedAttack.java:58-87) */
    MRTS_POLICY_CRUSH_V2 = 17, /* ai.abstraction.cRush.CRush_V2 (src/ai/abstraction/cRush/CRush_V2.java:69-479, CRanged_Tactic.java:77-389) */
    MRTS_POLICY_EMR_DETERMINISTICO = 18 /* ai.abstraction.EMRDeterministico (src/ai/abstraction/EMRDeterministico.java:74-358) */
};
enum { MRTS_PF_ASTAR = 0, MRTS_PF_BFS = 1, MRTS_PF_GREEDY = 2 /* ai.abstraction.pathfinding.GreedyPathFinding */,
       MRTS_PF_FLOODFILL = 3 /* ai.abstraction.pathfinding.FloodFillPathFinding: keeps its cache of distance maps per game and player in HBM */ };

/* Action row formats (8 int32 per row). */
enum {
    /* PlayerAction.fromVectorAction rows (src/rts/PlayerAction.java:384-417, src/rts/UnitAction.java:675-709):
     * [cell = x + y*W, type, moveDir, harvestDir, returnDir, produceDir, produceType, attackRelIdx] */
    MRTS_ACTIONS_VECTOR = 0,
    /* a PlayerAction as (unit, UnitAction) pairs: [cell of the unit, type, parameter, x, y, unitType(-1 none), 0, 0] */
    MRTS_ACTIONS_RAW = 1
};

/* MRTS_DTYPE_U8 keeps the low 8 bits of every value: hit points and resources above 255 (custom unit type tables, large resource
 * piles) wrap; use MRTS_DTYPE_I32 when the map or table can exceed that */
enum { MRTS_DTYPE_U8 = 0, MRTS_DTYPE_I32 = 1, MRTS_DTYPE_BITS = 2 /* masks only: element j of a row is bit j & 7 of byte j >> 3 */ };
enum {
    MRTS_FLAG_PARTIAL_OBS = 1u,
    MRTS_FLAG_SCRIPTED_AI = 2u, /* reserve pathfinding scratch so WORKER_RUSH / LIGHT_RUSH policies can be selected */
    MRTS_FLAG_PO_POLICIES = 4u  /* Game(partiallyObservable = true), src/rts/Game.java:129-134: each device policy decides on its
                                   player's PartiallyObservableGameState view; the lists go through issueSafe on the real state */
};
enum { MRTS_EVAL_SIMPLE_SQRT3 = 0, MRTS_EVAL_SIMPLE = 1 };

/* unit type fields for mrts_utt_get (order of the attributes in UnitType.toxml, src/rts/units/UnitType.java) */
enum {
    MRTS_UT_COST = 0, MRTS_UT_HP, MRTS_UT_MIN_DAMAGE, MRTS_UT_MAX_DAMAGE, MRTS_UT_ATTACK_RANGE, MRTS_UT_PRODUCE_TIME,
    MRTS_UT_MOVE_TIME, MRTS_UT_ATTACK_TIME, MRTS_UT_HARVEST_TIME, MRTS_UT_RETURN_TIME, MRTS_UT_HARVEST_AMOUNT,
    MRTS_UT_SIGHT_RADIUS, MRTS_UT_FLAGS /* isResource|isStockpile<<1|canHarvest<<2|canMove<<3|canAttack<<4 */,
    MRTS_UT_N_PRODUCES, MRTS_UT_PRODUCES0 /* .. +k */
};

/* per-game error bits (sticky) */
enum {
    MRTS_GE_UNIT_OVERFLOW = 1,      /* unit table capacity exceeded (a produce was dropped) */
    MRTS_GE_INCONSISTENT_OLDER = 2, /* GameState.issue "Inconsistent actions were executed!" branch (GameState.java:298-317) */
    MRTS_GE_FAILED_PRODUCE = 4,     /* produce completed without resources (UnitAction.java:457-461) */
    MRTS_GE_CELL_OCCUPIED = 8,      /* PhysicalGameState.addUnit would have thrown (PhysicalGameState.java:189-195) */
    MRTS_GE_BAD_ACTION = 16         /* malformed external action row */
};

typedef struct mrts_utt mrts_utt;
typedef struct mrts_map mrts_map;
typedef struct mrts_batch mrts_batch;

int mrts_abi_version(void);
const char *mrts_last_error(void);

/* ---- UnitTypeTable: new UnitTypeTable(version, crs), src/rts/units/UnitTypeTable.java:92-94,104-289 ---- */
int mrts_utt_create(int version, int conflict_policy, mrts_utt **out);
/* UnitTypeTable.fromJSON (src/rts/units/UnitTypeTable.java:393-411) */
int mrts_utt_from_json(const char *json, mrts_utt **out);
int mrts_utt_num_types(const mrts_utt *);
int mrts_utt_get(const mrts_utt *, int type_id, int field);
const char *mrts_utt_type_name(const mrts_utt *, int type_id);
int mrts_utt_conflict_policy(const mrts_utt *);
int mrts_utt_max_attack_range(const mrts_utt *);
void mrts_utt_destroy(mrts_utt *);

/* ---- PhysicalGameState.load(file, utt) / fromXML, src/rts/PhysicalGameState.java:65-76,700-726 ---- */
int mrts_map_load_xml(const char *path, const mrts_utt *, mrts_map **out);
int mrts_map_from_xml(const char *xml_text, const mrts_utt *, mrts_map **out);
/* programmatic construction (MapGenerator-style); units rows = [type, id, player, x, y, resources, hitpoints] */
int mrts_map_create(int width, int height, const uint8_t *terrain, int res0, int res1, int n_units,
                    const int32_t *units, const mrts_utt *, mrts_map **out);
int mrts_map_width(const mrts_map *);
int mrts_map_height(const mrts_map *);
int mrts_map_num_units(const mrts_map *);
int mrts_map_get_units(const mrts_map *, int32_t *out /* [n][7] */);
int mrts_map_get_terrain(const mrts_map *, uint8_t *out /* [h*w] */);
int mrts_map_resources(const mrts_map *, int player);
void mrts_map_destroy(mrts_map *);

/* ---- batch of n_games GameState objects on one device: new GameState(pgs, utt), GameState.java:62-65 ----
 * Game g starts from maps[g % n_maps]; all maps must share width and height. unit_capacity 0 = automatic
 * (initial units + total resources, rounded up; at most 252; a given capacity is rounded up to a multiple of 4). */
int mrts_batch_create(const mrts_utt *, const mrts_map *const *maps, int n_maps, int64_t n_games, int device,
                      uint32_t flags, int unit_capacity, mrts_batch **out);
void mrts_batch_destroy(mrts_batch *);
int64_t mrts_batch_num_games(const mrts_batch *);
int mrts_batch_unit_capacity(const mrts_batch *);
int mrts_batch_device(const mrts_batch *);
/* the CUDA stream (cudaStream_t) all work of this batch is enqueued on */
void *mrts_batch_stream(const mrts_batch *);
int mrts_batch_sync(mrts_batch *);

/* (Re)start every game from its map; seeds[g] seeds game g's java.util.Random streams (util/Sampler.java:17,
 * rts/GameState.java:37, rts/UnitAction.java:24).  seeds may be NULL (seed = game index). */
int mrts_batch_reset(mrts_batch *, const int64_t *seeds, int on_device);
/* restart only games whose mask byte is non-zero (auto-reset of finished environments,
 * src/tests/JNIGridnetVecClient.java:272-286) */
int mrts_batch_reset_masked(mrts_batch *, const uint8_t *mask, const int64_t *seeds, int on_device);

/* restart the masked games from their map like mrts_batch_reset_masked, but keep their RNG streams running (the reference's
 * Random objects are static and survive JNIGridnetVecClient's resets, src/tests/JNIGridnetVecClient.java:272-286) */
int mrts_batch_restart_masked(mrts_batch *, const uint8_t *mask, int on_device);

/* GameState.clone() (src/rts/GameState.java:582-604; NaiveMCTS clones the leaf's state before every playout, NaiveMCTS.java:
 * 201) for many games at once: game g of dst becomes a copy of game src_index[g] of src (src_index NULL: game g), for the games
 * whose mask byte is non-zero (mask NULL: all).  Both batches must share map size, unit capacity, flags' unit words and device.
 * The counters of dst (mrts_batch_stats) are not changed. */
int mrts_batch_copy_games(mrts_batch *dst, const mrts_batch *src, const int64_t *src_index /* [dst games] or NULL */,
                          const uint8_t *mask /* [dst games] or NULL */, int on_device);

int mrts_batch_set_policy(mrts_batch *, int player, int policy, int pathfinder);
/* Order of the two players inside one mrts_batch_step cycle.  0 (default): both PlayerActions are built on the pre-issue
 * state, then issueSafe(p0), issueSafe(p1) -- Game.start / JNIGridnetClient.gameStep (src/rts/Game.java:134-137,
 * src/tests/JNIGridnetClient.java:168-179).  1: player 1's PlayerAction is built on the state that already holds player
 * 0's -- JNIGridnetClientSelfPlay.gameStep (src/tests/JNIGridnetClientSelfPlay.java:160-170). */
int mrts_batch_set_issue_order(mrts_batch *, int sequential);
/* Step facts for the reward functions (src/ai/reward, all classes), written by every later mrts_batch_step into
 * out = [n_games][2 players][12] (device pointer; NULL disables), accumulated over the decision points of the step:
 *   [0] HARVEST and [1] RETURN actions in the player's PlayerAction as issueSafe left it (ResourceGatherRewardFunction),
 *   [2] ATTACKs on a cell held by the opponent, [3] on an own unit (AttackRewardFunction),
 *   [4] PRODUCE Worker, [5] PRODUCE Barracks/Base, [6] PRODUCE Light/Heavy/Ranged (Produce*RewardFunction),
 *   [7] the opponent had a Base before the cycle; [8]/[9] squared distance from it to the player's closest
 *       Worker/Light/Heavy/Ranged before / after the cycle, -1 when there is none (CloserToEnemyBaseRewardFunction),
 *   [10] some Resource unit still holds resources after the cycle, [11] internal.
 * WinLossRewardFunction reads mrts_batch_results. */
int mrts_batch_set_info_output(mrts_batch *, int32_t *out);
/* When enabled, mrts_batch_step restarts a game from its map at the start of the step if the game ended (game over or
 * time >= max_cycles) in an earlier step -- the auto-reset of src/tests/JNIGridnetVecClient.java:272-286, done on the
 * device.  The game's RNG streams keep running across episodes, like the reference's static Random objects. */
int mrts_batch_set_auto_reset(mrts_batch *, int enable);

/* Stage one PlayerAction per game for `player`, consumed by the next mrts_batch_step when that player's policy
 * is EXTERNAL (decode of both players happens on the pre-issue state, as in JNIGridnetClientSelfPlay.gameStep).
 * actions = [n_games][max_k][8], counts = [n_games].  fill_none_duration: JNIAI pads idle units with NONE(1);
 * pass a negative value for no padding. */
int mrts_batch_set_actions(mrts_batch *, int player, int format, const int32_t *actions, const int32_t *counts,
                           int max_k, int fill_none_duration, int on_device);

/* GameState.issueSafe(pa) / GameState.issue(pa) right now (GameState.java:338-408 / 249-328), no cycle. */
/* Both players' PlayerActions of every game from ONE array in that same environment order: row block 2g = player 0 of game g, row
 * block 2g + 1 = player 1 ([2 * n_games][max_k][8], every game max_k rows).  One host -> device copy; async != 0 returns without
 * waiting for it (the host array, ideally pinned, must stay untouched until the batch is synchronised). */
int mrts_batch_set_actions_interleaved(mrts_batch *, int format, const int32_t *actions, int max_k, int fill_none_duration, int on_device, int async);
/* ... followed by the one-cycle step: JNIGridnetVecClient.gameStep of the self-play environments as ONE call (no cycle cap; the
 * environments' step limit is mrts_batch_set_vec_autoreset's) */
int mrts_batch_vec_step(mrts_batch *, const int32_t *actions /* [2 * n_games][max_k][8] */, int max_k, int on_device, int async);
int mrts_batch_issue(mrts_batch *, int player, int format, const int32_t *actions, const int32_t *counts,
                     int max_k, int fill_none_duration, int safe, int on_device);

/* Game.start loop body (src/rts/Game.java:126-140) for up to n_cycles cycles per game:
 *   pa0 = policy0(gs); pa1 = policy1(gs); gs.issueSafe(pa0); gs.issueSafe(pa1); gameover = gs.cycle();
 * A game stops at game over or when time >= max_cycles. */
int mrts_batch_step(mrts_batch *, int n_cycles, int max_cycles);
/* GameState.cycle() only (no policies): advance every unfinished game up to absolute time t_target[g]
 * (or by n_cycles when t_target is NULL). TestTracesIntegrity.java:81-85 */
int mrts_batch_cycle_to(mrts_batch *, const int32_t *t_target, int n_cycles, int on_device);

/* EvaluationFunction.evaluate(maxplayer, 1 - maxplayer, gs) of every game's current state (src/ai/evaluation/
 * EvaluationFunction.java; eval_fn 0 = SimpleSqrtEvaluationFunction3.java:24-44, 1 = SimpleEvaluationFunction.java:21-36), of
 * observer's PartiallyObservableGameState view when observer >= 0.  out_eval = [n_games] float.  The batch is not modified. */
int mrts_batch_evaluate(mrts_batch *, int eval_fn, int maxplayer, int observer, float *out_eval, int on_device);

/* PathFinding.findPathToPositionInRange(start, targetpos, range, gs, null) for one unit per game
 * (src/ai/abstraction/pathfinding/PathFinding.java:17-24; AStarPathFinding.java:52-79, BFSPathFinding.java:41-147,
 * GreedyPathFinding.java:53-84): queries = [n_games][3] {cell of the start unit (x + y*W), target position (x + y*W), range};
 * range < 0 asks for PathFinding.findPath.  out_dir[g] = direction of the returned MOVE (0 up, 1 right, 2 down, 3 left) or -1
 * for null (no path, already in range, no unit on the cell).  The batch must own pathfinding scratch (MRTS_FLAG_SCRIPTED_AI);
 * it is not modified. */
int mrts_batch_pathfind(mrts_batch *, int pathfinder, const int32_t *queries, int32_t *out_dir, int on_device);

/* NaiveMCTS.simulate + evaluate (src/ai/mcts/naivemcts/NaiveMCTS.java:195-223,297-308): for every game g and every
 * k < rollouts_per_game, clone the game's state (gs2 = leaf.gs.clone(), from observer's PartiallyObservableGameState
 * when observer >= 0, src/rts/PartiallyObservableGameState.java:35-79), play RandomBiasedAI vs itself with issue() until
 * game over or `depth` cycles, and evaluate for maxplayer.  The batch itself is not modified.
 * Rollout r = g*rollouts_per_game + k is seeded with seeds[r] (r when seeds is NULL).
 * out_eval[r] = ef.evaluate(maxplayer, 1-maxplayer, gs2) as float (SimpleSqrtEvaluationFunction3.java:24-44 or
 * SimpleEvaluationFunction.java:21-36); out_time[r] = gs2.getTime() - start.  The caller applies the reference's
 * discount evaluation * Math.pow(0.99, time/10.0) in double precision on the host (NaiveMCTS.java:205). */
int mrts_batch_rollout(mrts_batch *, int rollouts_per_game, int depth, int eval_fn, int maxplayer, int observer,
                       const int64_t *seeds, float *out_eval, int32_t *out_time, int on_device);

/* GameState.getVectorObservation(player) (GameState.java:922-968; PartiallyObservableGameState.java:82-154 when the
 * batch has MRTS_FLAG_PARTIAL_OBS): out = [n_games][C][H][W], C = 6 or 8.  player < 0: per-game player array
 * not supported yet. */
int mrts_batch_observe(mrts_batch *, int player, int dtype, void *out, int on_device);
int mrts_batch_num_planes(const mrts_batch *);
/* Fused emission: while set (device pointers on the batch's device, 16-byte aligned; NULL disables a player), every
 * mrts_batch_step also writes GameState.getVectorObservation(player) of the state it leaves behind into
 * out_playerP = [n_games][C][H][W] -- what JNIGridnetVecClient.gameStep returns per environment
 * (src/tests/JNIGridnetVecClient.java:213-297) -- from the state that is already in shared memory, without a second pass
 * over the batch. */
int mrts_batch_set_observation_outputs(mrts_batch *, int dtype, void *out_player0, void *out_player1);
/* The same for the action masks (JNIGridnetClient.getMasks, src/tests/JNIGridnetClient.java:210-223; UnitAction.getValidActionArray,
 * src/rts/UnitAction.java:711-751): every later mrts_batch_step also writes the bit-packed masks of the state it leaves behind,
 * [n_games][H][W][(mask width + 7) / 8] bytes per player (element j of a cell in bit j & 7 of byte j >> 3, as MRTS_DTYPE_BITS of
 * mrts_batch_masks), into 16-byte aligned device buffers; NULL disables a player.  One launch then carries the step, both
 * observations and both masks. */
int mrts_batch_set_mask_outputs(mrts_batch *, void *out_player0, void *out_player1);
/* JNIGridnetVecClient.gameStep's auto-reset (src/tests/JNIGridnetVecClient.java:244-262,272-286) inside the step launch: after the
 * cycle, a game whose environment is done -- done_mode 1: the game is over (WinLossRewardFunction first); 2: no Resource unit holds
 * resources (ResourceGatherRewardFunction first); 3: never by state -- or that has taken max_steps steps since its last reset, is
 * restarted from its map before the fused observations / masks are written (the RNG streams keep running).  The step's results and
 * reward facts stay the terminal ones; bit 1 of results[g][2] reports the restart.  done_mode 0 switches it off (the default). */
int mrts_batch_set_vec_autoreset(mrts_batch *, int done_mode, int max_steps);
/* Game g's fused outputs (observations, masks) are written at game slot g * game_stride of their buffers (default 1).  With stride 2,
 * out_player0 = base and out_player1 = base + one game's bytes, the two players of game g land next to each other: the environment
 * order of JNIGridnetVecClient's self-play pairs (src/tests/JNIGridnetVecClient.java:226-236, environment 2g = player 0 of game g). */
int mrts_batch_set_output_stride(mrts_batch *, int game_stride); /* mrts_batch_masks with on_device != 0 follows it too */
/* JNIGridnetClient.getMasks(player) (src/tests/JNIGridnetClient.java:210-223, UnitAction.java:711-751):
 * out = [n_games][H][W][mask_width]; with MRTS_DTYPE_BITS the last dimension is (mask_width + 7) / 8 bytes of packed bits
 * (79 entries -> 10 bytes per cell: 1/32 of the int32 array the reference allocates). */
int mrts_batch_masks(mrts_batch *, int player, int dtype, void *out, int on_device);
int mrts_batch_mask_width(const mrts_batch *);

/* Flat host copy of game states (parity checks, MCTS leaf import).  Arrays are sized by the caller:
 *  header  [count][8]    : time, res0, res1, n_units, winner(-1 none), gameover(0/1), error bits, next unit id
 *  units   [count][cap][8]: type, player, x, y, resources, hitpoints, id, has_action(0/1)   (list order)
 *  actions [count][cap][8]: type, parameter, x, y, unitType, issue time, order (rank in assignment insertion order), 0
 *  rng     [count][3]    : raw 48-bit states of the policy / conflict / damage streams                              */
typedef struct {
    int32_t *header;
    int32_t *units;
    int32_t *actions;
    int64_t *rng;
} mrts_state_host;
int mrts_batch_export(mrts_batch *, int64_t first, int64_t count, mrts_state_host *out);
int mrts_batch_import(mrts_batch *, int64_t first, int64_t count, const mrts_state_host *in);
/* Unit.getUnitActions (src/rts/units/Unit.java:382-522) of every idle unit of `player`, as ORDERED lists -- the order
 * RandomBiasedAI samples from and PlayerActionGenerator enumerates (src/rts/PlayerActionGenerator.java:56-106) -- together with what the
 * generator's constructor derives from the state (the resource usage of the assignments in flight, GameState.java:652-664):
 *   out_hdr[g][8]       choices; resources used by player 0, 1; resources of player 0, 1; positions used; time;
 *                       bit 0 gameover | (winner + 1) << 1 | bit 3 / bit 4: player 0 / 1 has a unit without an assignment
 *   out_positions[g][]  (unit capacity entries) linear positions x + y * W reserved by in-flight MOVE / PRODUCE, unit-list order
 *   out_choices[g][c]   {unit slot, Unit.ID, type | x << 8 | y << 16 | (owner + 1) << 24, number of actions} for the c-th idle unit in
 *                       unit-list order (at most max_choices are written; out_hdr[g][0] is the full count)
 *   out_lists[g][c][k]  the k-th action of that unit: type | (direction + 1) << 4 | x << 8 | y << 16 | (unit type + 1) << 24
 *                       (x, y: the attacked cell; a NONE action lasts none_duration; lists longer than max_actions are cut)
 * GameState.getPlayerActions / PlayerActionGenerator stay host-side on top of this (microrts_b200/csrc/player_actions.hpp). */
int mrts_batch_unit_actions(mrts_batch *, int player, int none_duration, int max_choices, int max_actions, int32_t *out_hdr,
                            int32_t *out_positions, int32_t *out_choices, int32_t *out_lists, int on_device);
/* cycle() while the game is not over and neither player has a unit without an assignment -- the loop at the head of every MCTS node
 * (src/ai/mcts/naivemcts/NaiveMCTSNode.java:48-53) */
int mrts_batch_cycle_to_decision(mrts_batch *);
/* ---- host side of the search AIs, over the calls above -------------------------------------------------------------------------
 * GameState.getPlayerActions(player) (src/rts/GameState.java:493-524, PlayerAction.cartesianProduct src/rts/PlayerAction.java:180-195)
 * of one game, in the reference's order: PlayerAction i = out_counts[i] RAW rows at out_rows[i][..][8] ({cell, type, parameter, x, y,
 * unit type, 0, 0}: what mrts_batch_issue takes).  At most max_player_actions are written; *out_total is the full count. */
int mrts_batch_player_actions(mrts_batch *, int64_t game, int player, int32_t *out_rows /* [max_player_actions][max_k][8] */,
                              int32_t *out_counts, int64_t max_player_actions, int max_k, int64_t *out_total);
/* rts.PlayerActionGenerator (src/rts/PlayerActionGenerator.java:56-252) for one game and player.  The reference's static / unseeded
 * generators become a java.util.Random state the caller owns (mrts_java_random_seed(seed) = the state of new Random(seed)). */
typedef struct mrts_pag mrts_pag;
#define MRTS_PAG_DONE (-100) /* mrts_pag_next: getNextAction returned null (every PlayerAction has been generated) */
int mrts_pag_create(mrts_batch *, int64_t game, int player, int none_duration, mrts_pag **out); /* MRTS_E_STATE: no unit can act */
void mrts_pag_destroy(mrts_pag *);
int64_t mrts_pag_size(const mrts_pag *);       /* getSize(): product of the list sizes, capped at Long.MAX_VALUE */
int64_t mrts_pag_generated(const mrts_pag *);  /* getGenerated() */
int mrts_pag_num_choices(const mrts_pag *);
int mrts_pag_next(mrts_pag *, int32_t *rows /* [max_k][8] RAW */, int max_k);   /* getNextAction(-1): number of rows, or MRTS_PAG_DONE */
int mrts_pag_random(mrts_pag *, int64_t *rng_state, int32_t *rows, int max_k);  /* getRandom() */
int mrts_pag_randomize_order(mrts_pag *, int64_t *rng_state);                   /* randomizeOrder() */
int64_t mrts_java_random_seed(int64_t seed);

/* ai.mcts.naivemcts.NaiveMCTS (src/ai/mcts/naivemcts/NaiveMCTS.java:140-158,195-262; NaiveMCTSNode.java) or ai.mcts.uct.UCT for every game of `roots` at once:
 * search t looks for `player`'s best PlayerAction in game t.  The trees live on the host; node states, cloneIssue, the nodes' cycle
 * loops, the move generators' lists and the playouts (RandomBiasedAI both sides, `lookahead` cycles, eval_fn) run on the device for all
 * searches in lockstep.  Search t's generators are seeded from seeds[t] (the reference's are unseeded statics). */
typedef struct mrts_mcts mrts_mcts;
typedef struct {
    int lookahead;          /* MAXSIMULATIONTIME (100) */
    int max_depth;          /* MAX_TREE_DEPTH (10) */
    float epsilon_l, epsilon_g, epsilon_0; /* 0.3, 0.0, 0.4 */
    int global_strategy;    /* 0 = E_GREEDY, 1 = UCB1 */
    int force_exploration;  /* forceExplorationOfNonSampledActions */
    int eval_fn;            /* 0 = SimpleSqrtEvaluationFunction3, 1 = SimpleEvaluationFunction */
    int algorithm;          /* 0 = NaiveMCTS; 1 = UCT (src/ai/mcts/uct/UCT.java:103-199, UCTNode.java: every node's shuffled PlayerActionGenerator
                             * is exhausted before UCB1 (C = 0.05) picks among the children; the epsilons and the strategy are unused) */
} mrts_mcts_params;
int mrts_mcts_create(mrts_batch *roots, int player, const mrts_mcts_params *, int max_nodes_per_tree, const int64_t *seeds, mrts_mcts **out);
int mrts_mcts_iterate(mrts_mcts *, int n_iterations);  /* NaiveMCTS.iteration n times per search */
int mrts_mcts_num_nodes(const mrts_mcts *, int64_t tree);
int mrts_mcts_root(const mrts_mcts *, int64_t tree, int32_t *root_visits, double *root_accum, int32_t *child_visits, double *child_accum, int max_children);
int mrts_mcts_best_actions(const mrts_mcts *, int32_t *out_rows /* [n_games][max_k][8] RAW */, int32_t *out_counts, int max_k); /* getBestActionSoFar */
void mrts_mcts_destroy(mrts_mcts *);
/* game s of src becomes game dst_index[s] of dst (entries < 0 are skipped): the scatter form of mrts_batch_copy_games */
int mrts_batch_scatter_games(mrts_batch *dst, const mrts_batch *src, const int64_t *dst_index /* [src games] */, int on_device);

/* light per-game result: out[g] = {time, winner(-1 none), gameover, error bits} */
int mrts_batch_results(mrts_batch *, int32_t *out /* [n_games][4] */, int on_device);

/* Queue a device -> host copy behind the batch's pending work on its stream and return at once (mrts_batch_sync waits): how a host
 * fetches the fused outputs (observations, masks, reward facts) into its own -- ideally pinned -- arrays without a second stream. */
int mrts_batch_copy_to_host(mrts_batch *, void *host_dst, const void *device_src, size_t bytes);
/* counters since the last reset: {wins_p0, wins_p1, draws, games_finished, cycles, decisions, unit_cycles, errors} */
int mrts_batch_stats(mrts_batch *, int64_t out[8]);
/* Symbol of the step kernel the batch launched last (which of the specialised / fixed-layout copies ran), for benchmark reports. */
const char *mrts_batch_last_kernel(const mrts_batch *);
/* Bytes of game state and outputs the batch's kernels have read from (out[0]) and written to (out[1]) global memory since the last
 * full reset: headers + live unit words per load / store, observation planes, masks, results.  What the roofline's "real traffic"
 * figure is checked against (map templates, a few KB per map and L2 resident, are not counted). */
int mrts_batch_io_bytes(mrts_batch *, int64_t out[2]);

/* The run's single collective: the eight counters summed over every rank's batch with one ncclAllReduce(int64,
 * sum) on the batch's stream; every rank receives the totals.  comm NULL = mrts_batch_stats.  The reference has no counterpart
 * (it is single-process); a multi-GPU host -- rts.cuda.BatchedGameState with one thread or process per GPU -- creates the
 * communicator once: one rank calls mrts_nccl_unique_id and hands the 128 bytes to the others (a Java array between threads, a file or
 * socket between processes), then every rank calls mrts_nccl_comm_create (collective).  mrts_nccl_comm_wrap adopts an existing
 * ncclComm_t.  NCCL is resolved with dlopen("libnccl.so.2") at the first call: single-GPU users need no NCCL at all. */
#define MRTS_NCCL_UNIQUE_ID_BYTES 128
typedef struct mrts_comm mrts_comm;
int mrts_nccl_unique_id(uint8_t out[MRTS_NCCL_UNIQUE_ID_BYTES]);
int mrts_nccl_comm_create(const uint8_t id[MRTS_NCCL_UNIQUE_ID_BYTES], int n_ranks, int rank, int device, mrts_comm **out);
int mrts_nccl_comm_wrap(void *nccl_comm /* ncclComm_t */, int device, mrts_comm **out);
void mrts_nccl_comm_destroy(mrts_comm *);
int mrts_batch_stats_allreduce(mrts_batch *, int64_t out[8], mrts_comm *comm_or_null);
/* number of kernels this batch has launched so far */
int64_t mrts_batch_launch_count(const mrts_batch *);

#ifdef __cplusplus
}
#endif
#endif /* MICRORTS_CUDA_H */
