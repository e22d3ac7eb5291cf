"""Parity at the shapes bench.py actually launches (BASELINE.json configs 3, 4 and 5), against the CPU oracle.

  cfg 5  GardenOfWar64x64: fused step + observation planes (u8 and i32) and the action masks (dense and bit-packed),
         one cycle per step, both players, every step -- GameState.getVectorObservation (src/rts/GameState.java:922-968),
         UnitAction.getValidActionArray (src/rts/UnitAction.java:711-751), JNIGridnetClient.getMasks (src/tests/JNIGridnetClient.java:210-223)
  cfg 3  one batch over basesWorkers24x24{,A..L} round-robin, WorkerRush vs LightRush with A* in both side assignments,
         100 cycles per step, full games -- src/ai/abstraction/WorkerRush.java:63-204, LightRush.java:77-258
  cfg 4  BWDistantResources32x32: NaiveMCTS.simulate playouts (src/ai/mcts/naivemcts/NaiveMCTS.java:297-308) of depth 100 from the
         observer's partially observable view of CONTACT roots (an enemy unit in sight), >= 1024 rollouts
"""
import numpy as np
import pytest

import microrts_b200 as M
import parity as P
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _device_buffer(backend, shape, dtype):
    if backend == "emu":
        return np.full(shape, 0x55, dtype=dtype)
    import torch
    return torch.full(shape, 0x55, dtype=torch.uint8 if dtype == np.uint8 else torch.int32, device="cuda")


def _to_numpy(buf):
    return buf if isinstance(buf, np.ndarray) else buf.cpu().numpy()


@pytest.mark.parametrize("dtype", [np.uint8, np.int32])
def test_cfg5_fused_observations_and_masks_64x64(backend, maps, dtype):
    key = "GardenOfWar64x64"
    n = 2 if backend == "emu" else 64
    steps = 6 if backend == "emu" else 300
    m = maps[key]
    utt, outt = M.UnitTypeTable(1, 1), O.Utt(1, 1)
    b = M.BatchedGameState(utt, M.maps.standard_map(key, utt), n)
    seeds = np.arange(n, dtype=np.int64) * 104729 + 9
    b.reset(seeds)
    b.set_policy(0, M.POLICY_RANDOM_BIASED)
    b.set_policy(1, M.POLICY_RANDOM_BIASED)
    games = []
    for g in range(n):
        og = O.Game(outt, m)
        og.seed(int(seeds[g]))
        games.append(og)
    # spread the games over the phases of a match before the one-cycle steps start (game g is (g mod 8) * 180 cycles in)
    pre = 0 if backend == "emu" else 180
    sub = M.BatchedGameState(utt, M.maps.standard_map(key, utt), n)
    sub.set_policy(0, M.POLICY_RANDOM_BIASED)
    sub.set_policy(1, M.POLICY_RANDOM_BIASED)
    for k in range(1, 8 if pre else 1):
        mask = (np.arange(n) % 8 >= k).astype(np.uint8)
        sub.copy_games(b)
        sub.step(pre, 3000)
        b.copy_games(sub, mask=mask)
        for g, og in enumerate(games):
            if mask[g]:
                og.run(O.AI_RANDOM_BIASED, None, O.AI_RANDOM_BIASED, None, pre, 3000)
    sub.close()
    ex = b.export()
    for g in range(n):
        P.assert_same_state(ex, g, games[g], "pre-advance")
    shape = (n, 6, m["h"], m["w"])
    o0, o1 = _device_buffer(backend, shape, dtype), _device_buffer(backend, shape, dtype)
    b.set_observation_outputs(o0, o1)
    mb = (b.mask_width + 7) // 8
    assert b.mask_width == 79
    # the configuration bench.py --workload obs --with-masks launches: one kernel per step writes both observations and both masks
    f0, f1 = _device_buffer(backend, (n, m["h"], m["w"], mb), np.uint8), _device_buffer(backend, (n, m["h"], m["w"], mb), np.uint8)
    b.set_mask_outputs(f0, f1)
    mask_games = list(range(n)) if backend == "emu" else list(range(0, n, 4))
    for it in range(steps):
        b.step(1, 3000)
        b.sync()
        a = [_to_numpy(o0), _to_numpy(o1)]
        check_masks = it % 10 == 0 or it == steps - 1
        launches = b.launch_count
        mk_bits = [b.masks(pl, "bits") for pl in (0, 1)]
        fused = [_to_numpy(f0), _to_numpy(f1)]
        for pl in (0, 1):
            assert (fused[pl] == mk_bits[pl]).all(), "fused masks differ from mrts_batch_masks at step %d player %d" % (it, pl)
        assert b.launch_count == launches + 2
        mk_dense = [b.masks(pl, np.uint8) for pl in (0, 1)] if check_masks else None
        for g, og in enumerate(games):
            if not (og.gameover and og.time > 0):
                og.run(O.AI_RANDOM_BIASED, None, O.AI_RANDOM_BIASED, None, 1, 3000)
            for pl in (0, 1):
                ref = og.observe(pl)
                assert (a[pl][g] == ref.astype(dtype)).all(), "fused observation %s step %d game %d player %d" % (key, it, g, pl)
            if check_masks or g in mask_games:
                for pl in (0, 1):
                    ref_m = og.masks(pl)
                    assert mk_bits[pl][g].shape == (m["h"], m["w"], mb)
                    assert (np.unpackbits(mk_bits[pl][g], axis=-1, bitorder="little")[..., :79] == ref_m).all(), "bit-packed masks step %d game %d player %d" % (it, g, pl)
                    if check_masks:
                        assert (mk_dense[pl][g] == ref_m.astype(np.uint8)).all(), "dense masks step %d game %d player %d" % (it, g, pl)
    ex = b.export()
    for g in range(n):
        P.assert_same_state(ex, g, games[g], "after the one-cycle steps")
    b.close()


CFG3_KEYS = ["24x24/basesWorkers24x24"] + ["24x24/basesWorkers24x24" + c for c in "ABCDEFGHIJKL"]


@pytest.mark.parametrize("sides", [("WORKER_RUSH", "LIGHT_RUSH"), ("LIGHT_RUSH", "WORKER_RUSH")])
def test_cfg3_variant_batch_worker_rush_vs_light_rush(backend, maps, sides):
    """The batch bench.py --workload scripted launches: game g plays on variant g mod 13, 100 cycles per step, 3000-cycle cap,
    through the rush-only fixed-layout kernel; full games state-for-state."""
    n = 13 if backend == "emu" else 52
    total = 300 if backend == "emu" else 3000
    utt, outt = M.UnitTypeTable(1, 1), O.Utt(1, 1)
    b = M.BatchedGameState(utt, [M.maps.standard_map(k, utt) for k in CFG3_KEYS], n, scripted_ai=True)
    kinds = []
    for pl, name in enumerate(sides):
        b.set_policy(pl, getattr(M, "POLICY_" + name), M.PF_ASTAR)
        kinds.append(getattr(O, "AI_" + name))
    games, ais = [], []
    for g in range(n):
        games.append(O.Game(outt, maps[CFG3_KEYS[g % 13]]))
        ais.append([O.ScriptedAI(k, O.PF_ASTAR) for k in kinds])
    finished = 0
    for t in range(0, total, 100):
        b.step(100, 3000)
        ex = b.export()
        for g, og in enumerate(games):
            if not (og.gameover and og.time > 0):
                over, _ = og.run(kinds[0], ais[g][0], kinds[1], ais[g][1], 100, 3000)
                finished += 1 if over else 0
            P.assert_same_state(ex, g, og, "cfg3 %s map %s t=%d" % ("/".join(sides), CFG3_KEYS[g % 13], t + 100))
    res = b.results()
    assert (res[:, 3] == 0).all()
    if backend != "emu":
        assert finished == n, "every WorkerRush vs LightRush game on the 24x24 variants ends before the cap"
    b.close()


def test_cfg4_contact_root_rollouts(backend, maps):
    """>= 1024 playouts from partially observable CONTACT roots: evaluation and simulated cycles equal the oracle's."""
    key = "BWDistantResources32x32"
    n = 2 if backend == "emu" else 16
    R = 3 if backend == "emu" else 64
    depth, observer = 100, 0
    utt, outt = M.UnitTypeTable(1, 1), O.Utt(1, 1)
    pgs = M.maps.standard_map(key, utt)
    play = M.BatchedGameState(utt, pgs, n)
    roots = M.BatchedGameState(utt, pgs, n)
    seeds = np.arange(n, dtype=np.int64) + 4242
    play.reset(seeds)
    play.set_policy(0, M.POLICY_RANDOM_BIASED)
    play.set_policy(1, M.POLICY_RANDOM_BIASED)
    games = []
    for g in range(n):
        og = O.Game(outt, maps[key])
        og.seed(int(seeds[g]))
        games.append(og)
    # advance every game to the first multiple of 50 cycles at which the observer sees an enemy unit; the device finds the same
    # moment through the evaluation of the observer's view (no visible enemy <=> SimpleSqrtEvaluationFunction3 == 1.0)
    frozen = np.zeros(n, dtype=bool)
    root_games = [None] * n
    for t in range(50, 3001, 50):
        play.step(50, 3000)
        ev = play.evaluate(0, observer, observer)
        res = play.results()
        contact = (ev != 1.0) & ~frozen & (res[:, 2] == 0)
        for g, og in enumerate(games):
            if frozen[g]:
                continue
            og.run(O.AI_RANDOM_BIASED, None, O.AI_RANDOM_BIASED, None, 50, 3000)
            seen = (og.po_view(observer).units()[:, 1] == 1 - observer).any()
            if not og.gameover:
                assert bool(contact[g]) == bool(seen and np.float32(og.po_view(observer).evaluate(0, observer, 1 - observer)) != 1.0), (g, t)
        if contact.any():
            roots.copy_games(play, mask=contact.astype(np.uint8))
            for g in np.nonzero(contact)[0]:
                root_games[g] = games[g].clone()
            frozen |= contact
        if frozen.all() or backend == "emu" and t >= 600:
            break
    have = [g for g in range(n) if root_games[g] is not None]
    if backend != "emu":
        assert len(have) >= n * 3 // 4, "most RandomBiased games on this map reach contact"
    ex = roots.export()
    for g in have:
        P.assert_same_state(ex, g, root_games[g], "contact root")
    rs = np.arange(n * R, dtype=np.int64) * 48271 + 11
    ev, tm = roots.rollout(depth=depth, rollouts_per_game=R, eval_fn=0, maxplayer=observer, observer=observer, seeds=rs)
    total_cycles = 0
    for g in have:
        for k in range(R):
            c = root_games[g].po_view(observer)
            c.seed(int(rs[g * R + k]))
            start = c.time
            c.simulate(start + depth)
            assert tm[g, k] == c.time - start, (g, k, tm[g, k], c.time - start)
            assert ev[g, k] == np.float32(c.evaluate(0, observer, 1 - observer)), (g, k)
            total_cycles += c.time - start
    if backend != "emu":
        assert len(have) * R >= 768 and total_cycles / (len(have) * R) > 10, "contact roots play on instead of ending at the first cycle"
    play.close()
    roots.close()


def test_cfg5_full_size_fused_outputs_equal_the_standalone_kernels(backend, maps):
    """65 536 games of GardenOfWar64x64 exactly as bench.py --workload obs --with-masks launches them: after every one-cycle step the
    planes and masks the fused kernel wrote (bulk stores + scattered values) equal, for EVERY game, what the standalone observation
    kernel and the mask mode of the generic kernel write from the same state (plain stores, independent code paths), and a sample of
    games equals the oracle.  Size-independent check of the store ordering (zeros before values) under full load."""
    if backend == "emu":
        pytest.skip("full-size batch runs on the GPU only")
    import torch
    key, n = "GardenOfWar64x64", 65536
    m = maps[key]
    utt, outt = M.UnitTypeTable(1, 1), O.Utt(1, 1)
    b = M.BatchedGameState(utt, M.maps.standard_map(key, utt), n)
    seeds = np.arange(n, dtype=np.int64) + 77
    b.reset(seeds)
    b.set_policy(0, M.POLICY_RANDOM_BIASED)
    b.set_policy(1, M.POLICY_RANDOM_BIASED)
    b.step(400, 3000)  # into the mid-game, without outputs
    mb = (b.mask_width + 7) // 8
    obs = [torch.full((n, 6, 64, 64), 0x55, dtype=torch.uint8, device="cuda") for _ in range(2)]
    msk = [torch.full((n, 64, 64, mb), 0x55, dtype=torch.uint8, device="cuda") for _ in range(2)]
    ref_o = torch.empty((n, 6, 64, 64), dtype=torch.uint8, device="cuda")
    ref_m = torch.empty((n, 64, 64, mb), dtype=torch.uint8, device="cuda")
    b.set_observation_outputs(obs[0], obs[1])
    b.set_mask_outputs(msk[0], msk[1])
    sample = list(range(0, 24)) + [4095, 32768, 65535]
    games = []
    for g in sample:
        og = O.Game(outt, m)
        og.seed(int(seeds[g]))
        og.run(O.AI_RANDOM_BIASED, None, O.AI_RANDOM_BIASED, None, 400, 3000)
        games.append(og)
    terrain = torch.tensor([int(c) for c in m["terrain"]], dtype=torch.uint8, device="cuda").reshape(64, 64)
    for it in range(6):
        b.step(1, 3000)
        assert b.last_kernel == "k_step_fast_obs"
        b.sync()
        for pl in (0, 1):
            b.observe(pl, np.uint8, out=ref_o)
            b.sync()
            assert torch.equal(obs[pl], ref_o), "fused planes of player %d differ from k_observe at step %d" % (pl, it)
            assert bool((obs[pl][:, 5] == terrain).all())
            b.masks(pl, "bits", out=ref_m)
            b.sync()
            assert torch.equal(msk[pl], ref_m), "fused masks of player %d differ from the mask kernel at step %d" % (pl, it)
        for og in games:
            og.run(O.AI_RANDOM_BIASED, None, O.AI_RANDOM_BIASED, None, 1, 3000)
        for pl in (0, 1):
            o_h, m_h = obs[pl][sample].cpu().numpy(), msk[pl][sample].cpu().numpy()
            for k, og in enumerate(games):
                assert (o_h[k] == og.observe(pl).astype(np.uint8)).all(), (it, sample[k], pl)
                assert (np.unpackbits(m_h[k], axis=-1, bitorder="little")[..., :79] == og.masks(pl)).all(), (it, sample[k], pl)
    assert (b.results()[:, 3] == 0).all()
    b.close()
