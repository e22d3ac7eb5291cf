"""The RL-style contract of the drop-in, measured: JNIGridnetVecClient.gameStep (src/tests/JNIGridnetVecClient.java:213-297) for a
batch of self-play environments, one game cycle per step -- vector actions in, observations + action masks + rewards + dones out.

Used by bench.py (`--workload vec` and the `secondary.vec` entry of the headline line).  Three legs over the same environments:

  value                 device-resident: actions already in HBM, one step launch per gameStep (decode, issueSafe x2, cycle, reward facts,
                        auto-reset, both observations, both bit-packed masks), outputs left in HBM -- what a policy network on the same
                        GPU consumes.  environment-steps/s (an environment is one player's seat: two per game).
  e2e                   through vec_client.JNIGridnetVecClient(compact=True) with HOST buffers: pinned int32 actions in; uint8 observations,
                        bit-packed masks, rewards and dones out, every step, inside the timed region.
  e2e_reference_layout  the same with the reference's own array types (int32 observations from gameStep + int32 masks from getMasks(0)):
                        87 KB per 16x16 environment and step, i.e. PCIe bound.
Actions are synthetic: every environment sends K = 16 random rows per step (most address cells without an own idle unit and are
ignored, as the reference ignores them); units that get no row are filled with NONE(1) as JNIAI does.
"""
import time

import numpy as np

K_ROWS = 16
MAX_STEPS = 2000


def _random_actions(rng, n_envs, cells):
    a = np.zeros((n_envs, K_ROWS, 8), dtype=np.int32)
    a[:, :, 0] = rng.integers(0, cells, size=(n_envs, K_ROWS))
    a[:, :, 1] = rng.integers(0, 6, size=(n_envs, K_ROWS))
    a[:, :, 2:6] = rng.integers(0, 4, size=(n_envs, K_ROWS, 4))
    a[:, :, 6] = rng.integers(1, 7, size=(n_envs, K_ROWS))
    a[:, :, 7] = rng.integers(0, 49, size=(n_envs, K_ROWS))
    return a


def run(ctx, n_envs, steps, warmup, prewarm, cpu_seconds, cpu_baseline, timed_window, roofline, key="16x16/basesWorkers16x16"):
    """timed_window / roofline / cpu_baseline: bench.py's own timing rules and report builders (bench.py is the entry point)."""
    M, torch = ctx.M, ctx.torch
    from . import rewards as R
    from .vec_client import JNIGridnetVecClient
    n_envs -= n_envs % 2
    n_games = n_envs // 2
    pgs = ctx.pgs(key)
    W, H = pgs.getWidth(), pgs.getHeight()
    rng = np.random.default_rng(1234 + ctx.rank)
    pool = [torch.from_numpy(_random_actions(rng, n_envs, W * H)).pin_memory() for _ in range(4)]
    rfs = [R.WinLossRewardFunction(), R.ResourceGatherRewardFunction(), R.ProduceWorkerRewardFunction(), R.ProduceBuildingRewardFunction(),
           R.AttackRewardFunction(), R.ProduceCombatUnitRewardFunction()]

    # ---- device-resident leg: the batch exactly as the client configures it, driven without host copies -------------------
    b = M.BatchedGameState(ctx.utt, pgs, n_games, device=ctx.local)
    b.set_policy(0, M.POLICY_EXTERNAL); b.set_policy(1, M.POLICY_EXTERNAL); b.set_issue_order(True)
    obs = torch.empty((n_envs, 6, H, W), dtype=torch.uint8, device="cuda")
    mb = (b.mask_width + 7) // 8
    msk = torch.empty((n_envs, H, W, mb), dtype=torch.uint8, device="cuda")
    info = torch.zeros((n_games, 2, 12), dtype=torch.int32, device="cuda")
    b.set_info_output(info)
    b.set_output_layout(obs, msk, interleaved=True)
    b.set_vec_autoreset(1, MAX_STEPS)
    b.reset(ctx.seeds(n_games))
    dev_pool = [p.cuda() for p in pool]
    it = [0]

    def dev_step():
        b.vec_step(dev_pool[it[0] % len(dev_pool)])
        it[0] += 1
    w = timed_window(ctx, b, dev_step, steps, warmup, prewarm)
    d = w["stats"]
    env_steps = steps * n_envs
    mean_units = d["unit_cycles"] / max(1, d["cycles"])
    bytes_alg = d["cycles"] * 2.0 * (32.0 + 24.0 * mean_units) + steps * n_envs * (6 * H * W + H * W * mb)
    out = dict(value=env_steps * ctx.world / w["wall"], unit="env-steps/s", game_cycles_per_sec=d["cycles"] * ctx.world / w["wall"],
               ms_per_step=1000.0 * w["wall"] / max(1, steps), clocks=w["clocks"], gpu_launches=w["launches"],
               config=dict(workload="JNIGridnetVecClient flow: %d self-play environments/GPU (%d games of maps/%s.xml), one cycle per gameStep, %d synthetic "
                                    "vector-action rows per environment in; observations, bit-packed masks, reward facts and results out" % (n_envs, n_games, key, K_ROWS),
                           envs_per_gpu=n_envs, rows_per_env=K_ROWS, max_steps=MAX_STEPS, mean_live_units=mean_units,
                           note="value: actions and outputs stay in HBM (uint8 observations + bit-packed masks); e2e: pinned host buffers both ways through vec_client"),
               roofline=roofline(ctx, w, bytes_alg, "game-cycles x 2*(32+24*U) + env-steps x (6*H*W + H*W*%d)" % mb, "vec",
                                 note="one k_step launch per gameStep: state in, one cycle, state + observations + masks out"),
               stats=dict(device_time_s=w["dev_max"], wall_time_s=w["wall"], window_game_cycles=d["cycles"]))
    b.close()

    # ---- end to end through the client, host buffers both ways --------------------------------------------------------------
    def client_leg(compact, with_dense_masks, steps, warmup):
        vc = JNIGridnetVecClient(n_envs, 0, MAX_STEPS, rfs, "", [pgs] * n_envs, [], ctx.utt, partial_obs=False, device=ctx.local, seed=ctx.rank * n_envs,
                                 compact=compact)
        players = [0] * n_envs
        vc.reset(players)
        k = 0
        for _ in range(warmup):
            vc.gameStep(pool[k % len(pool)].numpy(), players); k += 1
            if with_dense_masks:
                vc.getMasks(0)
        ctx.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            r = vc.gameStep(pool[k % len(pool)].numpy(), players); k += 1
            if with_dense_masks:
                vc.getMasks(0)
        ctx.barrier()
        dt = time.perf_counter() - t0
        (dt,) = ctx.max_over_ranks(dt)
        g = vc.groups[0]
        osz = 6 * H * W * (1 if compact else 4)
        msz = H * W * (mb if compact else 4 * g.b.mask_width)
        d2h = n_envs * (osz + msz) + n_games * (16 + 96)
        h2d = n_envs * K_ROWS * 32
        assert r.observation.shape[0] == n_envs
        vc.close()
        return dict(value=steps * n_envs * ctx.world / dt, unit="env-steps/s", ms_per_step=1000.0 * dt / steps, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                    pcie_gbs=(h2d + d2h) * steps / dt / 1e9)

    out["e2e"] = dict(client_leg(True, False, steps, warmup), how="vec_client.JNIGridnetVecClient(compact=True).gameStep: uint8 observations + bit-packed masks + rewards/dones to pinned host arrays")
    out["e2e_reference_layout"] = dict(client_leg(False, True, max(2, min(steps, 5)), 1), how="gameStep (int32 observations) + getMasks(0) (int32 masks), the reference's array types: PCIe bound")
    if cpu_seconds > 0 and ctx.rank == 0:
        cb = cpu_baseline("vec", key, cpu_seconds)
        cb["note"] = "game-cycles/s of the port with both observations per cycle; one game-cycle = 2 environment-steps"
        out["cpu_baseline"] = cb
    return out
