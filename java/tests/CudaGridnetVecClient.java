package tests;

import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.ValueLayout;
import java.nio.file.Path;

import ai.reward.RewardFunctionInterface;
import rts.cuda.BatchedGameState;
import rts.units.UnitTypeTable;
import tests.JNIGridnetVecClient.Responses;

/**
 * tests.JNIGridnetVecClient (src/tests/JNIGridnetVecClient.java:106-316) for self-play environments over libmicrorts_cuda.so: same
 * reset / gameStep / getMasks / close surface and array conventions (reused buffers owned by the client, environment 2g = player 0
 * of game g), one kernel launch per gameStep instead of one Java GameState per environment.
 *
 * What the native side does per gameStep (include/microrts_cuda.h): decode of both players' vector actions
 * (mrts_batch_set_actions_interleaved), issueSafe in self-play order, cycle, the reward facts, the auto-reset of finished
 * environments (mrts_batch_set_vec_autoreset) and the observations of every environment in this class's order
 * (mrts_batch_set_observation_outputs + mrts_batch_set_output_stride).  The reward values are computed here from the facts, with the
 * constants of ai.reward.* (the Python mirror microrts_b200/rewards.py does the same and is what the tests exercise).
 *
 * NOT COMPILED in the build container of this repository (no JDK there); the Python mirror microrts_b200/vec_client.py is the
 * tested implementation of exactly this flow.  Device memory comes from the caller (cudaMalloc through any CUDA binding): obsDev,
 * infoDev, resDev.  Only num_envs == 0 (self-play) is shown; agent-vs-bot environments add setPolicy(1 - side, POLICY_*)
 * on a batch created with FLAG_SCRIPTED_AI (plus FLAG_PO_POLICIES when partial_obs: the opponent then decides on its own view, as
 * ai2.getAction(1 - player, player2gs) does in JNIGridnetClient.gameStep).
 */
public class CudaGridnetVecClient {
    final BatchedGameState gs;
    final int numSelfPlayEnvs, maxSteps, planes, h, w, nRf;
    final Arena arena = Arena.ofConfined();
    final MemorySegment actionsHost, obsHost, infoHost, resHost, obsDev, infoDev, resDev;
    final int[][][][] observation;
    final double[][] reward;
    final boolean[][] done;
    final Responses responses;
    final int maxK;

    public CudaGridnetVecClient(int a_num_selfplayenvs, int a_max_steps, RewardFunctionInterface[] a_rfs, String libraryPath, String mapPath,
                                UnitTypeTable a_utt, int maxK, MemorySegment obsDev, MemorySegment infoDev, MemorySegment resDev) throws Throwable {
        numSelfPlayEnvs = a_num_selfplayenvs; maxSteps = a_max_steps; nRf = a_rfs.length; this.maxK = maxK;
        gs = new BatchedGameState(Path.of(libraryPath), UnitTypeTable.VERSION_ORIGINAL_FINETUNED, a_utt.getMoveConflictResolutionStrategy(), mapPath,
                                  a_num_selfplayenvs / 2, 0, 0);
        planes = gs.planes; h = gs.height; w = gs.width;
        gs.setPolicy(0, BatchedGameState.POLICY_EXTERNAL);
        gs.setPolicy(1, BatchedGameState.POLICY_EXTERNAL);
        gs.setSequentialIssue(true);                                  // JNIGridnetClientSelfPlay.gameStep :160-170
        this.obsDev = obsDev; this.infoDev = infoDev; this.resDev = resDev;
        long obsBytesPerEnv = 4L * planes * h * w;
        gs.setDeviceOutputs(infoDev, BatchedGameState.DTYPE_I32, MemorySegment.NULL, MemorySegment.NULL);
        gs.setInterleavedOutputs(BatchedGameState.DTYPE_I32, obsDev, obsBytesPerEnv, MemorySegment.NULL, 0);
        gs.setVecAutoreset(1, a_max_steps);                           // done[0] of WinLossRewardFunction = game over; :244-262,272-286
        actionsHost = arena.allocate(ValueLayout.JAVA_INT, (long) numSelfPlayEnvs * maxK * 8);
        obsHost = arena.allocate(ValueLayout.JAVA_INT, (long) numSelfPlayEnvs * planes * h * w);
        infoHost = arena.allocate(ValueLayout.JAVA_INT, (long) (numSelfPlayEnvs / 2) * 2 * 12);
        resHost = arena.allocate(ValueLayout.JAVA_INT, (long) (numSelfPlayEnvs / 2) * 4);
        observation = new int[numSelfPlayEnvs][planes][h][w];
        reward = new double[numSelfPlayEnvs][nRf];
        done = new boolean[numSelfPlayEnvs][nRf];
        responses = new Responses(null, null, null);
    }

    public Responses reset(int[] players) throws Throwable {
        gs.reset(null);
        int[] flat = new int[numSelfPlayEnvs / 2 * planes * h * w];
        for (int p = 0; p < 2; p++) {
            gs.getVectorObservation(p, flat);
            for (int g = 0; g < numSelfPlayEnvs / 2; g++) unflatten(flat, g, observation[2 * g + p]);
        }
        responses.set(observation, reward, done);
        return responses;
    }

    public Responses gameStep(int[][][] action, int[] players) throws Throwable {
        long o = 0;
        for (int e = 0; e < numSelfPlayEnvs; e++)
            for (int k = 0; k < maxK; k++)
                for (int j = 0; j < 8; j++) actionsHost.setAtIndex(ValueLayout.JAVA_INT, o++, k < action[e].length ? action[e][k][j] : 0);
        gs.setActionsInterleaved(actionsHost, maxK, true);
        gs.step(1, Integer.MAX_VALUE);                                 // one launch: decode, issueSafe x2, cycle, facts, auto-reset, observations
        gs.copyToHost(resHost, resDev, resHost.byteSize());
        gs.copyToHost(infoHost, infoDev, infoHost.byteSize());
        gs.copyToHost(obsHost, obsDev, obsHost.byteSize());
        gs.sync();
        int[] flat = obsHost.toArray(ValueLayout.JAVA_INT);
        for (int e = 0; e < numSelfPlayEnvs; e++) unflatten(flat, e, observation[e]);
        for (int g = 0; g < numSelfPlayEnvs / 2; g++) {
            int winner = resHost.getAtIndex(ValueLayout.JAVA_INT, 4L * g + 1), flags = resHost.getAtIndex(ValueLayout.JAVA_INT, 4L * g + 2);
            boolean over = (flags & 1) != 0, restarted = (flags & 2) != 0;
            for (int p = 0; p < 2; p++) {
                int e = 2 * g + p;
                long f = (2L * g + p) * 12;
                // ai.reward.*: WinLoss, ResourceGather (+1 per HARVEST / RETURN), ProduceWorker, ProduceBuilding, Attack, ProduceCombatUnit
                double[] r = {over ? (winner == p ? 1.0 : -1.0) : 0.0,
                              infoHost.getAtIndex(ValueLayout.JAVA_INT, f) + infoHost.getAtIndex(ValueLayout.JAVA_INT, f + 1),
                              infoHost.getAtIndex(ValueLayout.JAVA_INT, f + 4), infoHost.getAtIndex(ValueLayout.JAVA_INT, f + 5),
                              infoHost.getAtIndex(ValueLayout.JAVA_INT, f + 2) - infoHost.getAtIndex(ValueLayout.JAVA_INT, f + 3),
                              infoHost.getAtIndex(ValueLayout.JAVA_INT, f + 6)};
                for (int j = 0; j < nRf && j < r.length; j++) { reward[e][j] = r[j]; done[e][j] = j == 0 ? over : (j == 1 && infoHost.getAtIndex(ValueLayout.JAVA_INT, 2L * g * 12 + 10) == 0); }
                if (restarted) done[e][0] = true;
            }
        }
        responses.set(observation, reward, done);
        return responses;
    }

    public int[][][][] getMasks(int player) throws Throwable {
        int k = gs.maskW;
        int[][][][] out = new int[numSelfPlayEnvs][h][w][k];
        int[] flat = new int[numSelfPlayEnvs / 2 * h * w * k];
        for (int p = 0; p < 2; p++) {
            gs.getMasks(p, flat);
            for (int g = 0; g < numSelfPlayEnvs / 2; g++)
                for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) System.arraycopy(flat, ((g * h + y) * w + x) * k, out[2 * g + p][y][x], 0, k);
        }
        return out;
    }

    private void unflatten(int[] flat, int index, int[][][] dst) {
        int o = index * planes * h * w;
        for (int c = 0; c < planes; c++) for (int y = 0; y < h; y++) { System.arraycopy(flat, o, dst[c][y], 0, w); o += w; }
    }

    public void close() {
        gs.close();
        arena.close();
    }
}
