"""Parity of the CUDA engine (through the C ABI / microrts_b200) against the CPU oracle and the golden traces.

Every test here needs a device (`-m gpu`).  `MRTS_EMU=1` runs them against the emulated-warp debug build instead.
The bar is bit-exact: unit list (order, type, owner, position, resources, hit points, ids), player resources, time,
winner, and the in-flight assignments (action, issue time, insertion order).
"""
import os

import numpy as np
import pytest

import microrts_b200 as M
import parity as P
from oracle import oracle as O

pytestmark = pytest.mark.gpu


# MRTS_SEED_OFFSET shifts the seeds of every randomized differential test: a soak run repeats the suite with other games
SOAK = int(os.environ.get("MRTS_SEED_OFFSET", "0"))


def make_pgs(mapd, utt):
    return M.PhysicalGameState.fromXML(P.map_to_xml(mapd), utt)


# ------------------------------------------------------------------------------------------------------------------
# golden traces: TestTracesIntegrity protocol (test/microrts/TestTracesIntegrity.java:72-127) on the device
# ------------------------------------------------------------------------------------------------------------------
def test_trace_replay_on_device(backend, traces, maps):
    utt = M.UnitTypeTable(1, 1)
    groups = {}
    for t in traces:
        m = maps[t["mapkey"]]
        groups.setdefault((m["w"], m["h"]), []).append(t)
    if backend == "emu":  # keep the emulated run short: one size group in three
        groups = {k: v for i, (k, v) in enumerate(sorted(groups.items())) if i % 3 == 0}
    n_checked = 0
    for (w, h), ts in sorted(groups.items()):
        pgs = [make_pgs(maps[t["mapkey"]], utt) for t in ts]
        b = M.BatchedGameState(utt, pgs, len(ts))
        rounds = max(len(t["entries"]) for t in ts)
        for r in range(rounds):
            target = np.array([t["entries"][min(r, len(t["entries"]) - 1)]["time"] for t in ts], dtype=np.int32)
            b.cycle_to(target)
            ex = b.export()
            max_k = 1
            for g, t in enumerate(ts):
                if r < len(t["entries"]):
                    max_k = max(max_k, len(t["entries"][r]["actions"]))
            rows = [np.zeros((len(ts), max_k, 8), dtype=np.int32) for _ in range(2)]
            counts = [np.zeros(len(ts), dtype=np.int32) for _ in range(2)]
            for g, t in enumerate(ts):
                if r >= len(t["entries"]):
                    continue
                e = t["entries"][r]
                hdr, units, _a = P.export_game(ex, g)
                exp = np.array(e["units"], dtype=np.int32).reshape(-1, 6)
                assert hdr[0] == e["time"], (t["name"], r)
                assert units[:, :6].shape == exp.shape and (units[:, :6] == exp).all(), \
                    "%s entry %d time %d\ndev=\n%s\ntrace=\n%s" % (t["name"], r, e["time"], units[:, :6], exp)
                assert (hdr[1], hdr[2]) == tuple(e["res"]), (t["name"], r)
                assert hdr[6] == 0, "error bits %d in %s" % (hdr[6], t["name"])
                n_checked += 1
                for (ui, ty, par, x, y, ut) in e["actions"]:
                    pl = exp[ui][1]
                    rows[pl][g, counts[pl][g]] = [exp[ui][2] + exp[ui][3] * w, ty, par, x, y, ut, 0, 0]
                    counts[pl][g] += 1
            if counts[0].any() or counts[1].any():
                b.issueSafe(0, rows[0], counts[0])
                b.issueSafe(1, rows[1], counts[1])
        b.close()
    assert n_checked > 1000


# ------------------------------------------------------------------------------------------------------------------
# RandomBiasedAI self-play (Game.start loop) against the oracle
# ------------------------------------------------------------------------------------------------------------------
SELFPLAY = [
    # mapkey, games (emu, cuda), cycles (emu, cuda), chunk
    ("8x8/basesWorkers8x8", (6, 96), (3000, 3000), 1),
    ("8x8/basesWorkers8x8", (6, 96), (3000, 3000), 53),
    ("16x16/basesWorkers16x16", (3, 48), (1200, 3000), 100),
    ("24x24/basesWorkers24x24", (2, 24), (700, 3000), 250),
    ("BWDistantResources32x32", (2, 16), (600, 3000), 64),
    ("GardenOfWar64x64", (1, 8), (400, 3000), 500),
    ("8x8/FourBasesWorkers8x8", (4, 32), (1500, 3000), 7),
    ("melee14x12Mixed18", (2, 32), (600, 2000), 10),
]


def run_selfplay(backend, maps, key, sizes, cycles, chunk, version=1, conflict=1, check_every=1, seed0=1000):
    n = sizes[0] if backend == "emu" else sizes[1]
    total = cycles[0] if backend == "emu" else cycles[1]
    utt, outt = M.UnitTypeTable(version, conflict), O.Utt(version, conflict)
    b = M.BatchedGameState(utt, make_pgs(maps[key], utt), n)
    seeds = np.arange(n, dtype=np.int64) * 7919 + seed0 + SOAK
    b.reset(seeds)
    b.set_policy(0, M.POLICY_RANDOM_BIASED)
    b.set_policy(1, M.POLICY_RANDOM_BIASED)
    games = []
    for g in range(n):
        og = O.Game(outt, maps[key])
        og.seed(int(seeds[g]))
        games.append(og)
    it = 0
    for t in range(0, total, chunk):
        b.step(chunk, total)
        for og in games:
            if not (og.gameover and og.time > 0):
                og.run(O.AI_RANDOM_BIASED, None, O.AI_RANDOM_BIASED, None, chunk, total)
        it += 1
        if it % check_every == 0 or t + chunk >= total:
            ex = b.export()
            for g in range(n):
                P.assert_same_state(ex, g, games[g], "%s t=%d" % (key, t + chunk))
                assert [int(v) for v in ex["rng"][g]] == [games[g].rng_state(k) for k in range(3)], "RNG state differs"
    res = b.results()
    for g in range(n):
        assert res[g, 0] == games[g].time and res[g, 1] == games[g].winner and res[g, 2] == int(games[g].gameover)
        assert res[g, 3] == 0
    st = b.stats()
    assert st["games_finished"] == sum(1 for og in games if og.gameover or og.time >= total)
    assert st["wins_p0"] == sum(1 for og in games if og.gameover and og.winner == 0)
    assert st["wins_p1"] == sum(1 for og in games if og.gameover and og.winner == 1)
    assert st["cycles"] == sum(og.time for og in games)
    b.close()
    return games


def games_rng(og):
    return og.rng_state(0)


@pytest.mark.parametrize("key,sizes,cycles,chunk", SELFPLAY)
def test_random_biased_selfplay(backend, maps, key, sizes, cycles, chunk):
    run_selfplay(backend, maps, key, sizes, cycles, chunk, check_every=1 if chunk > 20 else 25)


@pytest.mark.parametrize("version,conflict", [(2, 1), (3, 1), (1, 2), (1, 3), (3, 2), (2, 3)])
def test_utt_versions_and_conflict_policies(backend, maps, version, conflict):
    run_selfplay(backend, maps, "8x8/basesWorkers8x8", (4, 64), (2000, 3000), 11, version=version, conflict=conflict, check_every=10)
    run_selfplay(backend, maps, "16x16/TwoBasesBarracks16x16", (2, 32), (500, 2000), 40, version=version, conflict=conflict)


# ------------------------------------------------------------------------------------------------------------------
# observation planes and action masks (GameState.getVectorObservation / JNIGridnetClient.getMasks)
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("key", ["8x8/basesWorkers8x8", "16x16/basesWorkers16x16", "melee14x12Mixed18", "BWDistantResources32x32"])
def test_observations_and_masks(backend, maps, key):
    n = 3 if backend == "emu" else 24
    chunks = 4 if backend == "emu" else 10
    utt, outt = M.UnitTypeTable(1, 1), O.Utt(1, 1)
    pgs = make_pgs(maps[key], utt)
    b = M.BatchedGameState(utt, pgs, n)
    bpo = M.BatchedGameState(utt, pgs, n, partial_obs=True)
    seeds = np.arange(n, dtype=np.int64) + 5 + SOAK
    games = []
    for bb in (b, bpo):
        bb.reset(seeds)
        bb.set_policy(0, M.POLICY_RANDOM_BIASED)
        bb.set_policy(1, M.POLICY_RANDOM_BIASED)
    for g in range(n):
        og = O.Game(outt, maps[key])
        og.seed(int(seeds[g]))
        games.append(og)
    for c in range(chunks):
        for player in (0, 1):
            obs_i = b.observe(player, np.int32)
            obs_b = b.observe(player, np.uint8)
            po = bpo.observe(player, np.int32)
            po_b = bpo.observe(player, np.uint8)
            mk = b.masks(player, np.int32)
            mk_b = b.masks(player, np.uint8)
            mk_bits = b.masks(player, "bits")
            assert obs_i.shape == (n, 6, maps[key]["h"], maps[key]["w"]) and po.shape[1] == 8
            for g, og in enumerate(games):
                ref = og.observe(player)
                assert (obs_i[g] == ref).all(), (key, c, g, player)
                assert (obs_b[g] == ref.astype(np.uint8)).all()
                ref_po = og.po_view(player).observe(player, po=True)
                assert (po[g] == ref_po).all(), "PO obs %s chunk %d game %d player %d\n%s\n%s" % (key, c, g, player, po[g], ref_po)
                assert (po_b[g] == ref_po.astype(np.uint8)).all()
                ref_m = og.masks(player)
                assert (mk[g] == ref_m).all(), "masks %s chunk %d game %d player %d" % (key, c, g, player)
                assert (mk_b[g] == ref_m.astype(np.uint8)).all()
                assert (np.unpackbits(mk_bits[g], axis=-1, bitorder="little")[..., :ref_m.shape[-1]] == ref_m).all(), "bit-packed masks"

        for bb in (b, bpo):
            bb.step(37, 3000)
        for og in games:
            og.run(O.AI_RANDOM_BIASED, None, O.AI_RANDOM_BIASED, None, 37, 3000)
    b.close()
    bpo.close()


def _device_buffer(backend, shape, dtype):
    """A buffer the engine can write 'on device': a torch CUDA tensor, or plain host memory under the emulator."""
    if backend == "emu":
        return np.full(shape, 0x55, dtype=dtype)
    import torch
    return torch.full(shape, 0x55, dtype=torch.uint8 if dtype == np.uint8 else torch.int32, device="cuda")


def _to_numpy(buf):
    return buf if isinstance(buf, np.ndarray) else buf.cpu().numpy()


@pytest.mark.parametrize("key,dtype,external", [("16x16/basesWorkers16x16", np.uint8, False), ("melee14x12Mixed18", np.int32, False),
                                                ("NoWhereToRun9x8", np.uint8, True), ("8x8/basesWorkers8x8", np.int32, True)])
def test_fused_step_observations(backend, maps, key, dtype, external):
    """mrts_batch_set_observation_outputs: the planes written by the step kernel itself equal the oracle's
    getVectorObservation of the post-step state for both players (fast RandomBiased kernel, and the generic kernel when a
    player is EXTERNAL)."""
    n = 3 if backend == "emu" else 40
    utt, outt = M.UnitTypeTable(1, 1), O.Utt(1, 1)
    m = maps[key]
    b = M.BatchedGameState(utt, make_pgs(m, utt), n)
    seeds = np.arange(n, dtype=np.int64) + 77 + SOAK
    b.reset(seeds)
    b.set_policy(0, M.POLICY_RANDOM_BIASED)
    b.set_policy(1, M.POLICY_EXTERNAL if external else M.POLICY_RANDOM_BIASED)
    shape = (n, 6, m["h"], m["w"])
    o0, o1 = _device_buffer(backend, shape, dtype), _device_buffer(backend, shape, dtype)
    b.set_observation_outputs(o0, o1)
    # ... and the fused bit-packed action masks (mrts_batch_set_mask_outputs)
    mshape = (n, m["h"], m["w"], (b.mask_width + 7) // 8)
    k0, k1 = _device_buffer(backend, mshape, np.uint8), _device_buffer(backend, mshape, np.uint8)
    b.set_mask_outputs(k0, k1)
    games = []
    for g in range(n):
        og = O.Game(outt, m)
        og.seed(int(seeds[g]))
        games.append(og)
    for it in range(6):
        b.step(23, 3000)
        b.sync()
        a0, a1 = _to_numpy(o0), _to_numpy(o1)
        mk = [_to_numpy(k0), _to_numpy(k1)]
        for g, og in enumerate(games):
            og.run(O.AI_RANDOM_BIASED, None, O.AI_PASSIVE if external else O.AI_RANDOM_BIASED, None, 23, 3000)
            assert (a0[g] == og.observe(0).astype(dtype)).all(), (key, it, g)
            assert (a1[g] == og.observe(1).astype(dtype)).all(), (key, it, g)
            for pl in (0, 1):
                ref_m = og.masks(pl)
                assert (np.unpackbits(mk[pl][g], axis=-1, bitorder="little")[..., :ref_m.shape[-1]] == ref_m).all(), "fused masks %s it %d game %d player %d" % (key, it, g, pl)
    b.set_mask_outputs(None, None)
    # player 1 only, then disabled: untouched buffers stay untouched
    b.set_observation_outputs(None, o1)
    before = _to_numpy(o0).copy()
    b.step(5, 3000)
    b.sync()
    assert (_to_numpy(o0) == before).all()
    b.set_observation_outputs(None, None)
    b.close()


# ------------------------------------------------------------------------------------------------------------------
# vector actions through the EXTERNAL policy (JNIGridnetClientSelfPlay.gameStep flow) against the oracle
# ------------------------------------------------------------------------------------------------------------------
def random_vector_actions(rng, og, player, w, h, max_k):
    """A mix of sensible and nonsense rows: mostly rows addressing own idle units with a random legal-looking action."""
    units = og.units()
    asg = og.assignments()
    rows = []
    for i, u in enumerate(units):
        if u[1] != player or asg[i][0]:
            continue
        if rng.random() < 0.15:
            continue
        acts = og.unit_actions(i)
        if rng.random() < 0.8:
            ty, par, x, y, ut = acts[rng.integers(len(acts))]
        else:  # possibly illegal
            ty, par, x, y, ut = int(rng.integers(6)), int(rng.integers(4)), int(u[2] + rng.integers(-2, 3)), int(u[3] + rng.integers(-2, 3)), int(rng.integers(1, 7))
        row = [int(u[2] + u[3] * w), ty, 0, 0, 0, 0, 0, 0]
        if ty == O.MOVE: row[2] = par
        elif ty == O.HARVEST: row[3] = par
        elif ty == O.RETURN: row[4] = par
        elif ty == O.PRODUCE: row[5], row[6] = par, ut
        elif ty == O.ATTACK: row[7] = int((y - u[3] + 3) * 7 + (x - u[2] + 3)) if abs(x - u[2]) <= 3 and abs(y - u[3]) <= 3 else 24
        rows.append(row)
    if rng.random() < 0.3:  # rows addressing empty cells / enemy units / busy units
        rows.append([int(rng.integers(w * h)), int(rng.integers(6)), 1, 1, 1, 1, 3, 10])
    rng.shuffle(rows)
    return rows[:max_k]


EXTERNAL_CASES = [("8x8/basesWorkers8x8", 1, 1), ("16x16/basesWorkers16x16", 1, 1),
                  # every unit type, walls, other UnitTypeTable versions and conflict policies
                  ("melee14x12Mixed18", 2, 1), ("BWDistantResources32x32", 3, 1), ("8x8/FourBasesWorkers8x8", 1, 2),
                  ("16x16/TwoBasesBarracks16x16", 1, 3), ("12x12/complexBasesWorkers12x12", 3, 2), ("16x16/melee16x16Mixed12", 2, 3)]


@pytest.mark.parametrize("key,version,conflict", EXTERNAL_CASES)
def test_external_vector_actions(backend, maps, key, version, conflict):
    n = 3 if backend == "emu" else 32
    total = 250 if backend == "emu" else 1200
    w, h = maps[key]["w"], maps[key]["h"]
    utt, outt = M.UnitTypeTable(version, conflict), O.Utt(version, conflict)
    b = M.BatchedGameState(utt, make_pgs(maps[key], utt), n)
    seeds = np.arange(n, dtype=np.int64) + 5 + SOAK  # random damage (v3) and CANCEL_RANDOM draw from the per-game streams
    b.reset(seeds)
    b.set_policy(0, M.POLICY_EXTERNAL)
    b.set_policy(1, M.POLICY_EXTERNAL)
    games = []
    for g in range(n):
        og = O.Game(outt, maps[key])
        og.seed(int(seeds[g]))
        games.append(og)
    rng = np.random.default_rng(7)
    max_k = 24
    for t in range(total):
        rows = [np.zeros((n, max_k, 8), dtype=np.int32) for _ in range(2)]
        counts = [np.zeros(n, dtype=np.int32) for _ in range(2)]
        pas = []
        for g, og in enumerate(games):
            pa = []
            for player in (0, 1):
                r = random_vector_actions(rng, og, player, w, h, max_k)
                counts[player][g] = len(r)
                if r:
                    rows[player][g, :len(r)] = r
                pa.append(og.from_vector_action(player, np.array(r, dtype=np.int32).reshape(-1, 8), fill_none=1))
            pas.append(pa)
        for player in (0, 1):
            b.set_actions(player, rows[player], counts[player], M.ACTIONS_VECTOR, fill_none_duration=1)
        b.step(1, 5000)
        ex = b.export()
        for g, og in enumerate(games):
            if not (og.gameover and og.time > 0):
                og.issue(pas[g][0], True)
                og.issue(pas[g][1], True)
                og.cycle()
            P.assert_same_state(ex, g, og, "%s vector t=%d" % (key, t))
    b.close()


def test_auto_reset_and_masked_reset(backend, maps):
    key = "8x8/basesWorkers8x8"
    n = 4 if backend == "emu" else 64
    cap = 400  # short cap so that episodes roll over
    utt, outt = M.UnitTypeTable(1, 1), O.Utt(1, 1)
    b = M.BatchedGameState(utt, make_pgs(maps[key], utt), n)
    seeds = np.arange(n, dtype=np.int64) + 77 + SOAK
    b.reset(seeds)
    b.set_policy(0, M.POLICY_RANDOM_BIASED)
    b.set_policy(1, M.POLICY_RANDOM_BIASED)
    b.set_auto_reset(True)
    games = []
    for g in range(n):
        og = O.Game(outt, maps[key])
        og.seed(int(seeds[g]))
        games.append(og)
    for it in range(12):
        b.step(150, cap)
        for g in range(n):
            og = games[g]
            if (og.gameover and og.time > 0) or og.time >= cap:  # restart, RNG streams keep running
                ng = O.Game(outt, maps[key])
                for k in range(3):
                    ng.set_rng_state(k, og.rng_state(k))
                games[g] = og = ng
            og.run(O.AI_RANDOM_BIASED, None, O.AI_RANDOM_BIASED, None, 150, cap)
        ex = b.export()
        for g in range(n):
            P.assert_same_state(ex, g, games[g], "auto-reset it=%d" % it)
    # masked reset with fresh seeds
    b.set_auto_reset(False)
    mask = np.zeros(n, dtype=np.uint8)
    mask[::2] = 1
    new_seeds = seeds + 1000
    b.reset_masked(mask, new_seeds)
    ex = b.export()
    for g in range(n):
        if mask[g]:
            og = O.Game(outt, maps[key])
            og.seed(int(new_seeds[g]))
            games[g] = og
        P.assert_same_state(ex, g, games[g], "masked reset")
    b.close()


def test_export_import_roundtrip(backend, maps):
    key = "16x16/basesWorkers16x16"
    n = 3 if backend == "emu" else 32
    utt = M.UnitTypeTable(1, 1)
    pgs = make_pgs(maps[key], utt)
    a, c = M.BatchedGameState(utt, pgs, n), M.BatchedGameState(utt, pgs, n)
    for bb in (a, c):
        bb.set_policy(0, M.POLICY_RANDOM_BIASED)
        bb.set_policy(1, M.POLICY_RANDOM_BIASED)
    a.reset(np.arange(n, dtype=np.int64) + 9)
    a.step(333, 3000)
    st = a.export()
    c.import_(st)
    a.step(200, 3000)
    c.step(200, 3000)
    ea, ec = a.export(), c.export()
    for k in ("header", "units", "actions", "rng"):
        assert (ea[k] == ec[k]).all(), k
    a.close()
    c.close()


# ------------------------------------------------------------------------------------------------------------------
# BASELINE.json configs[1] at full size: properties that do not need the oracle at scale
# ------------------------------------------------------------------------------------------------------------------
def test_full_size_batch_properties(backend, maps):
    if backend == "emu":
        pytest.skip("full-size batch runs on the GPU only")
    key, n = "16x16/basesWorkers16x16", 65536
    utt, outt = M.UnitTypeTable(1, 1), O.Utt(1, 1)
    pgs = make_pgs(maps[key], utt)
    seeds = np.arange(n, dtype=np.int64) + SOAK
    b = M.BatchedGameState(utt, pgs, n)
    b.set_policy(0, M.POLICY_RANDOM_BIASED)
    b.set_policy(1, M.POLICY_RANDOM_BIASED)
    b.reset(seeds)
    for _ in range(6):
        b.step(500, 3000)
    res = b.results()
    st = b.stats()
    assert (res[:, 3] == 0).all(), "per-game error bits set"
    over = res[:, 2] != 0
    assert ((res[:, 0] == 3000) | over).all()
    assert st["games_finished"] == n and st["cycles"] == int(res[:, 0].sum())
    assert st["wins_p0"] == int(((res[:, 1] == 0) & over).sum()) and st["wins_p1"] == int(((res[:, 1] == 1) & over).sum())
    full = b.export(0, 4096)
    # determinism and independence of the batch partition: a sub-batch with the same seeds reproduces its games exactly
    sub = M.BatchedGameState(utt, pgs, 4096)
    sub.set_policy(0, M.POLICY_RANDOM_BIASED)
    sub.set_policy(1, M.POLICY_RANDOM_BIASED)
    sub.reset(seeds[:4096])
    sub.step(3000, 3000)
    es = sub.export()
    for k in ("header", "units", "actions", "rng"):
        assert (full[k] == es[k]).all(), k
    # spot-check a sample of the full batch against the oracle (whole games)
    for g in list(range(0, 64)) + [4095, 65535]:
        og = O.Game(outt, maps[key])
        og.seed(int(seeds[g]))
        og.run(O.AI_RANDOM_BIASED, None, O.AI_RANDOM_BIASED, None, 3000, 3000)
        ex = b.export(g, 1)
        P.assert_same_state(ex, 0, og, "full-size game %d" % g)
    b.close()
    sub.close()


# ------------------------------------------------------------------------------------------------------------------
# NaiveMCTS.simulate rollouts + evaluation functions (fully and partially observable roots)
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("key,observer", [("8x8/basesWorkers8x8", -1), ("8x8/basesWorkers8x8", 0), ("16x16/basesWorkers16x16", 1),
                                          ("BWDistantResources32x32", 0), ("melee14x12Mixed18", -1)])
def test_rollouts(backend, maps, key, observer):
    n = 2 if backend == "emu" else 16
    R = 2 if backend == "emu" else 6
    depth = 100
    utt, outt = M.UnitTypeTable(1, 1), O.Utt(1, 1)
    b = M.BatchedGameState(utt, make_pgs(maps[key], utt), n)
    seeds = np.arange(n, dtype=np.int64) + 3 + SOAK
    b.reset(seeds)
    b.set_policy(0, M.POLICY_RANDOM_BIASED)
    b.set_policy(1, M.POLICY_RANDOM_BIASED)
    games = []
    for g in range(n):
        og = O.Game(outt, maps[key])
        og.seed(int(seeds[g]))
        games.append(og)
    for warm in (0, 230, 500):  # roots: the initial state and two mid-game states
        if warm:
            b.step(warm, 3000)
            for og in games:
                og.run(O.AI_RANDOM_BIASED, None, O.AI_RANDOM_BIASED, None, warm, 3000)
        before = b.export()
        rs = np.arange(n * R, dtype=np.int64) * 31 + 17 + warm
        for fn in (0, 1):
            ev, tm = b.rollout(depth=depth, rollouts_per_game=R, eval_fn=fn, maxplayer=1, observer=observer, seeds=rs)
            for g, og in enumerate(games):
                for k in range(R):
                    c = og.po_view(observer) if observer >= 0 else og.clone()
                    c.seed(int(rs[g * R + k]))
                    start = c.time
                    c.simulate(start + depth)
                    ref = np.float32(c.evaluate(fn, 1, 0))
                    assert tm[g, k] == c.time - start, (key, warm, g, k, tm[g, k], c.time - start)
                    assert ev[g, k] == ref, (key, warm, fn, g, k, ev[g, k], ref)
        after = b.export()
        for kk in ("header", "units", "actions", "rng"):
            assert (before[kk] == after[kk]).all(), "rollout modified the batch"
    b.close()


# ------------------------------------------------------------------------------------------------------------------
# scripted policies on the device: LightRush / WorkerRush + AbstractionLayerAI + A* / BFS
# ------------------------------------------------------------------------------------------------------------------
def test_lightrush_traces_regenerated_on_device(backend, traces, maps):
    """Pinned by golden data: LightRush(A*) vs LightRush(A*) on the device must end every one of the 140 recorded mirror
    matches (src/tests/GenerateTestTraces.java:101-134, 500-cycle cap) in exactly the recorded final state."""
    utt = M.UnitTypeTable(1, 1)
    lr = [t for t in traces if "LightRush" in t["name"]]
    groups = {}
    for t in lr:
        m = maps[t["mapkey"]]
        groups.setdefault((m["w"], m["h"]), []).append(t)
    if backend == "emu":
        groups = {k: v for i, (k, v) in enumerate(sorted(groups.items())) if i % 5 == 0 and k[0] * k[1] <= 600}
    checked = 0
    for (w, h), ts in sorted(groups.items()):
        b = M.BatchedGameState(utt, [make_pgs(maps[t["mapkey"]], utt) for t in ts], len(ts), scripted_ai=True)
        b.set_policy(0, M.POLICY_LIGHT_RUSH)
        b.set_policy(1, M.POLICY_LIGHT_RUSH)
        b.step(500, 500)
        ex = b.export()
        for g, t in enumerate(ts):
            last = t["entries"][-1]
            hdr, units, _a = P.export_game(ex, g)
            exp = np.array(last["units"], dtype=np.int32).reshape(-1, 6)
            assert hdr[0] == last["time"], (t["name"], hdr[0], last["time"])
            assert units[:, :6].shape == exp.shape and (units[:, :6] == exp).all(), "%s\ndev=\n%s\ntrace=\n%s" % (t["name"], units[:, :6], exp)
            assert (hdr[1], hdr[2]) == tuple(last["res"]) and hdr[6] == 0, t["name"]
            checked += 1
        b.close()
    assert checked == (140 if backend != "emu" else checked)


SCRIPTED = [
    # map, policy0, policy1, pathfinder
    ("8x8/basesWorkers8x8", "LIGHT_RUSH", "WORKER_RUSH", 0),
    ("16x16/basesWorkers16x16", "WORKER_RUSH", "LIGHT_RUSH", 0),
    ("24x24/basesWorkers24x24", "WORKER_RUSH", "LIGHT_RUSH", 0),
    ("24x24/basesWorkers24x24H", "LIGHT_RUSH", "WORKER_RUSH", 1),
    ("16x16/TwoBasesBarracks16x16", "WORKER_RUSH", "WORKER_RUSH", 1),
    ("8x8/FourBasesWorkers8x8", "LIGHT_RUSH", "RANDOM_BIASED", 0),
    ("BWDistantResources32x32", "RANDOM_BIASED", "WORKER_RUSH", 0),
    ("16x16/basesWorkers16x16", "HEAVY_RUSH", "RANGED_RUSH", 0),
    ("8x8/basesWorkers8x8", "RANGED_RUSH", "LIGHT_RUSH", 1),
    ("24x24/basesWorkers24x24", "RANGED_RUSH", "HEAVY_RUSH", 0),
    # the defenses (WorkerDefense / LightDefense / HeavyDefense / RangedDefense) and GreedyPathFinding (pathfinder 2)
    ("8x8/basesWorkers8x8", "WORKER_DEFENSE", "WORKER_RUSH", 0),
    ("16x16/basesWorkers16x16", "LIGHT_RUSH", "LIGHT_DEFENSE", 0),
    ("16x16/basesWorkers16x16", "HEAVY_DEFENSE", "RANGED_DEFENSE", 1),
    ("24x24/basesWorkers24x24", "RANGED_DEFENSE", "WORKER_DEFENSE", 0),
    ("8x8/FourBasesWorkers8x8", "LIGHT_DEFENSE", "RANDOM_BIASED", 0),
    ("16x16/TwoBasesBarracks16x16", "RANDOM_BIASED", "HEAVY_DEFENSE", 2),
    ("16x16/basesWorkers16x16", "WORKER_RUSH", "LIGHT_RUSH", 2),
    ("8x8/basesWorkers8x8", "RANGED_RUSH", "WORKER_DEFENSE", 2),
    ("BWDistantResources32x32", "LIGHT_RUSH", "RANGED_RUSH", 2),
    ("16x16/basesWorkers16x16", "WORKER_RUSH_PP", "LIGHT_RUSH", 0),
    ("8x8/basesWorkers8x8", "WORKER_DEFENSE", "WORKER_RUSH_PP", 1),
    # FloodFillPathFinding (pathfinder 3): distance maps cached per target position in each AI instance across cycles
    ("8x8/basesWorkers8x8", "WORKER_RUSH", "LIGHT_RUSH", 3),
    ("16x16/basesWorkers16x16", "LIGHT_RUSH", "WORKER_RUSH", 3),
    ("16x16/TwoBasesBarracks16x16", "RANGED_RUSH", "HEAVY_RUSH", 3),
    ("24x24/basesWorkers24x24", "WORKER_RUSH", "LIGHT_DEFENSE", 3),
    ("BWDistantResources32x32", "LIGHT_RUSH", "WORKER_RUSH", 3),
    # CRush_V1 (ai/abstraction/cRush): worker rush on maps of at most 144 cells, barracks + kiting ranged units (RangedAttack)
    # on larger ones; Heavy units are the slower enemies its ranged units step back from
    ("8x8/basesWorkers8x8", "CRUSH_V1", "WORKER_RUSH", 0),
    ("8x8/FourBasesWorkers8x8", "LIGHT_RUSH", "CRUSH_V1", 0),
    ("16x16/basesWorkers16x16", "CRUSH_V1", "HEAVY_RUSH", 0),
    ("16x16/basesWorkers16x16", "LIGHT_RUSH", "CRUSH_V1", 1),
    ("16x16/TwoBasesBarracks16x16", "CRUSH_V1", "CRUSH_V1", 0),
    ("24x24/basesWorkers24x24", "HEAVY_RUSH", "CRUSH_V1", 0),
    ("BWDistantResources32x32", "CRUSH_V1", "HEAVY_DEFENSE", 3),
    ("16x16/basesWorkers16x16", "RANDOM_BIASED", "CRUSH_V1", 2),
    # CRush_V2: CRush_V1 whose combat units follow CRanged_Tactic from cycle 400 on (formation behind the leading Ranged unit,
    # attack when the enemy has nothing but workers left) -- long games
    ("8x8/basesWorkers8x8", "WORKER_RUSH", "CRUSH_V2", 0),
    ("16x16/basesWorkers16x16", "CRUSH_V2", "LIGHT_RUSH", 0),
    ("16x16/basesWorkers16x16", "HEAVY_RUSH", "CRUSH_V2", 0),
    ("16x16/basesWorkers16x16", "CRUSH_V2", "CRUSH_V2", 1),
    ("16x16/TwoBasesBarracks16x16", "CRUSH_V2", "RANDOM_BIASED", 0),
    ("24x24/basesWorkers24x24", "CRUSH_V2", "HEAVY_RUSH", 0),
    ("24x24/basesWorkers24x24", "CRUSH_V1", "CRUSH_V2", 0),
    ("BWDistantResources32x32", "RANGED_RUSH", "CRUSH_V2", 3),
    ("16x16/basesWorkers16x16", "CRUSH_V2", "LIGHT_DEFENSE", 2),
    # EMRDeterministico: economy first (up to 6 workers per base, more barracks and bases), Light / Ranged / Heavy in turn
    ("8x8/basesWorkers8x8", "EMR_DETERMINISTICO", "WORKER_RUSH", 0),
    ("16x16/basesWorkers16x16", "LIGHT_RUSH", "EMR_DETERMINISTICO", 0),
    ("24x24/basesWorkers24x24", "EMR_DETERMINISTICO", "HEAVY_RUSH", 1),
    ("BWDistantResources32x32", "EMR_DETERMINISTICO", "EMR_DETERMINISTICO", 0),
    ("BWDistantResources32x32", "PASSIVE", "EMR_DETERMINISTICO", 0),
    ("16x16/TwoBasesBarracks16x16", "EMR_DETERMINISTICO", "CRUSH_V2", 0),
    ("16x16/basesWorkers16x16", "EMR_DETERMINISTICO", "RANGED_DEFENSE", 3),
]


@pytest.mark.parametrize("key,p0,p1,pf", SCRIPTED)
def test_scripted_policies_vs_oracle(backend, maps, key, p0, p1, pf):
    n = 2 if backend == "emu" else 8
    total = 600 if backend == "emu" else 3000
    chunk = 25
    utt, outt = M.UnitTypeTable(1, 1), O.Utt(1, 1)
    b = M.BatchedGameState(utt, make_pgs(maps[key], utt), n, scripted_ai=True)
    seeds = np.arange(n, dtype=np.int64) + 11 + SOAK
    b.reset(seeds)
    kinds = []
    for pl, name in enumerate((p0, p1)):
        b.set_policy(pl, getattr(M, "POLICY_" + name), pf)
        kinds.append(getattr(O, "AI_" + name))
    games, ais = [], []
    for g in range(n):
        og = O.Game(outt, maps[key])
        og.seed(int(seeds[g]))
        games.append(og)
        ais.append([O.ScriptedAI(k, pf) if k in O.SCRIPTED_AIS else None for k in kinds])
    for t in range(0, total, chunk):
        b.step(chunk, total)
        ex = b.export()
        for g, og in enumerate(games):
            if not (og.gameover and og.time > 0):
                og.run(kinds[0], ais[g][0], kinds[1], ais[g][1], chunk, total)
            P.assert_same_state(ex, g, og, "%s %s/%s t=%d" % (key, p0, p1, t + chunk))
    b.close()


def test_emr_second_base_hashset_order_on_device(backend):
    """EMRDeterministico builds its second base next to the first element of a HashSet<Unit> of far resources
    (EMRDeterministico.java:264-311): hash order of the unit IDs, not list order.  Custom map of tests/test_oracle_golden.py
    (IDs 104 and 112: the resource listed last comes first), against the oracle and against the hand-derived outcome."""
    from test_oracle_golden import EMR_SECOND_BASE, _tiny_map
    mapd = dict(_tiny_map(EMR_SECOND_BASE, 16, 16), players=[[0, 15], [1, 5]])
    utt, outt = M.UnitTypeTable(1, 1), O.Utt(1, 1)
    for opp in ("PASSIVE", "WORKER_RUSH"):
        b = M.BatchedGameState(utt, make_pgs(mapd, utt), 4, scripted_ai=True)
        b.reset(np.arange(4, dtype=np.int64) + 3)
        b.set_policy(0, M.POLICY_EMR_DETERMINISTICO, 0)
        b.set_policy(1, getattr(M, "POLICY_" + opp), 0)
        og = O.Game(outt, mapd)
        og.seed(3)
        k1 = getattr(O, "AI_" + opp)
        a0, a1 = O.ScriptedAI(O.AI_EMR_DETERMINISTICO, 0), (O.ScriptedAI(k1, 0) if k1 in O.SCRIPTED_AIS else None)
        for t in range(0, 800, 40):
            b.step(40, 3000)
            if not (og.gameover and og.time > 0):
                og.run(O.AI_EMR_DETERMINISTICO, a0, k1, a1, 40, 3000)
            P.assert_same_state(b.export(), 0, og, "EMR second base vs %s t=%d" % (opp, t + 40))
            if opp == "PASSIVE" and t + 40 == 400:
                bases = sorted((int(u[2]), int(u[3])) for u in og.units() if u[0] == 1 and u[1] == 0)
                assert len(bases) >= 2 and all(q == (1, 1) or q[1] >= 12 for q in bases), bases
        b.close()


# ------------------------------------------------------------------------------------------------------------------
# Game(partiallyObservable = true), rts/Game.java:129-140: every device policy decides on its player's
# PartiallyObservableGameState view (units out of sight hidden, their reservations unknown), issueSafe on the real state;
# PO*Rush explore when they see no enemy (ai/abstraction/partialobservability/*.java)
# ------------------------------------------------------------------------------------------------------------------
PO_GAMES = [
    ("16x16/basesWorkers16x16", "PO_LIGHT_RUSH", "PO_WORKER_RUSH", 0),
    ("24x24/basesWorkers24x24", "PO_WORKER_RUSH", "PO_HEAVY_RUSH", 0),
    ("BWDistantResources32x32", "PO_RANGED_RUSH", "PO_LIGHT_RUSH", 0),
    ("16x16/basesWorkers16x16", "PO_HEAVY_RUSH", "LIGHT_DEFENSE", 1),
    ("8x8/basesWorkers8x8", "RANDOM_BIASED", "PO_WORKER_RUSH", 0),
    ("16x16/basesWorkers16x16", "RANDOM_BIASED", "RANDOM_BIASED", 0),
    ("16x16/TwoBasesBarracks16x16", "PO_LIGHT_RUSH", "RANDOM_BIASED", 2),
    ("24x24/basesWorkers24x24", "LIGHT_RUSH", "WORKER_RUSH", 0),
]


@pytest.mark.parametrize("key,p0,p1,pf", PO_GAMES)
def test_partially_observable_games_vs_oracle(backend, maps, key, p0, p1, pf):
    n = 2 if backend == "emu" else 8
    total = 600 if backend == "emu" else 3000
    chunk = 25
    utt, outt = M.UnitTypeTable(1, 1), O.Utt(1, 1)
    b = M.BatchedGameState(utt, make_pgs(maps[key], utt), n, scripted_ai=True, po_policies=True)
    seeds = np.arange(n, dtype=np.int64) + 23 + SOAK
    b.reset(seeds)
    kinds = []
    for pl, name in enumerate((p0, p1)):
        b.set_policy(pl, getattr(M, "POLICY_" + name), pf)
        kinds.append(getattr(O, "AI_" + name))
    games, ais = [], []
    for g in range(n):
        og = O.Game(outt, maps[key])
        og.seed(int(seeds[g]))
        games.append(og)
        ais.append([O.ScriptedAI(k, pf) if k in O.SCRIPTED_AIS else None for k in kinds])
    for t in range(0, total, chunk):
        b.step(chunk, total)
        ex = b.export()
        for g, og in enumerate(games):
            if not (og.gameover and og.time > 0):
                og.run_po(kinds[0], ais[g][0], kinds[1], ais[g][1], chunk, total)
            P.assert_same_state(ex, g, og, "PO %s %s/%s t=%d" % (key, p0, p1, t + chunk))
    b.close()


# ------------------------------------------------------------------------------------------------------------------
# wide differential: thousands of complete games, final state against the oracle (rare paths: more than 32 units per
# game, cross-chunk arbitration, cancel-both pairs, deaths with pending actions)
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("key,n_cuda", [("8x8/basesWorkers8x8", 4096), ("16x16/basesWorkers16x16", 2048), ("8x8/FourBasesWorkers8x8", 1024)])
def test_wide_differential_full_games(backend, maps, key, n_cuda):
    import threading
    n = 6 if backend == "emu" else n_cuda
    total = 600 if backend == "emu" else 3000
    utt, outt = M.UnitTypeTable(1, 1), O.Utt(1, 1)
    b = M.BatchedGameState(utt, make_pgs(maps[key], utt), n)
    seeds = np.arange(n, dtype=np.int64) * 3 + 17 + SOAK
    b.reset(seeds)
    b.set_policy(0, M.POLICY_RANDOM_BIASED)
    b.set_policy(1, M.POLICY_RANDOM_BIASED)
    b.step(total, total)
    ex = b.export()
    games = [None] * n
    nthreads = min(16, os.cpu_count() or 1)

    def work(t):
        for g in range(t, n, nthreads):
            og = O.Game(outt, maps[key])
            og.seed(int(seeds[g]))
            og.run(O.AI_RANDOM_BIASED, None, O.AI_RANDOM_BIASED, None, total, total)
            games[g] = og

    ts = [threading.Thread(target=work, args=(t,)) for t in range(nthreads)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    big = 0
    for g, og in enumerate(games):
        P.assert_same_state(ex, g, og, "%s wide game %d" % (key, g))
        big += og.n_units > 32
    assert (b.results()[:, 3] == 0).all()
    if backend != "emu" and key.startswith("16x16"):
        assert big > 100, "the batch should contain many games with more than 32 live units"
    b.close()


# ------------------------------------------------------------------------------------------------------------------
# every map of the reference (maps/**/*.xml as packed in tests/golden): RandomBiasedAI self-play on the fast kernel, and one
# of eight policy / pathfinder / observability combinations on the generic kernel (rotating over the maps), state for state
# against the oracle
# ------------------------------------------------------------------------------------------------------------------
MAP_SWEEP_COMBOS = [
    # policy 0, policy 1, pathfinder, partially observable game
    ("WORKER_DEFENSE", "LIGHT_RUSH", 0, False),
    ("RANGED_RUSH", "HEAVY_DEFENSE", 1, False),
    ("PO_LIGHT_RUSH", "PO_WORKER_RUSH", 0, True),
    ("LIGHT_DEFENSE", "RANDOM_BIASED", 2, False),
    ("PO_RANGED_RUSH", "RANDOM_BIASED", 1, True),
    ("WORKER_RUSH", "RANGED_DEFENSE", 0, False),
    ("RANDOM_BIASED", "RANDOM_BIASED", 0, True),
    ("PO_HEAVY_RUSH", "LIGHT_DEFENSE", 2, True),
    ("WORKER_RUSH_PP", "HEAVY_RUSH", 0, False),
    ("CRUSH_V1", "HEAVY_RUSH", 0, False),
    ("RANDOM_BIASED", "CRUSH_V1", 0, True),
    ("CRUSH_V2", "LIGHT_RUSH", 0, False),
    ("CRUSH_V2", "RANDOM_BIASED", 3, False),
    ("EMR_DETERMINISTICO", "WORKER_RUSH", 0, False),
]


@pytest.mark.parametrize("part", range(4))
def test_every_reference_map(backend, maps, part):
    keys = sorted(maps.keys())
    todo = list(enumerate(keys))[part::4]
    if backend == "emu":
        todo = todo[::12]
    n, total = (1, 100) if backend == "emu" else (4, 1500)
    utt, outt = M.UnitTypeTable(1, 1), O.Utt(1, 1)
    for idx, key in todo:
        big = maps[key]["w"] * maps[key]["h"] > 64 * 64
        for scripted in (False, True):
            if scripted:
                p0, p1, pf, po = MAP_SWEEP_COMBOS[idx % len(MAP_SWEEP_COMBOS)]
                cyc = min(total, 500 if big else 1000)  # pathfinding on the 128x128 maps: keep the oracle side short
            else:
                p0, p1, pf, po = "RANDOM_BIASED", "RANDOM_BIASED", 0, False
                cyc = total
            b = M.BatchedGameState(utt, make_pgs(maps[key], utt), n, scripted_ai=scripted, po_policies=po)
            seeds = np.arange(n, dtype=np.int64) * 7 + 5 + SOAK
            b.reset(seeds)
            kinds = [getattr(O, "AI_" + p0), getattr(O, "AI_" + p1)]
            b.set_policy(0, getattr(M, "POLICY_" + p0), pf)
            b.set_policy(1, getattr(M, "POLICY_" + p1), pf)
            b.step(cyc, cyc)
            ex = b.export()
            for g in range(n if not scripted else min(n, 2)):
                og = O.Game(outt, maps[key])
                og.seed(int(seeds[g]))
                ais = [O.ScriptedAI(k, pf) if k in O.SCRIPTED_AIS else None for k in kinds]
                (og.run_po if po else og.run)(kinds[0], ais[0], kinds[1], ais[1], cyc, cyc)
                P.assert_same_state(ex, g, og, "%s %s/%s pf=%d po=%s" % (key, p0, p1, pf, po))
            b.close()


# ------------------------------------------------------------------------------------------------------------------
# wide differential for the scripted policies: they are deterministic, so each runs against RandomBiasedAI opponents
# with hundreds of different seeds (complete games, final state against the oracle); half of the combinations as
# partially observable games
# ------------------------------------------------------------------------------------------------------------------
WIDE_SCRIPTED = [
    # map, scripted policy, its side, pathfinder, partially observable
    ("16x16/basesWorkers16x16", "LIGHT_RUSH", 0, 0, False),
    ("16x16/basesWorkers16x16", "WORKER_RUSH", 1, 1, False),
    ("8x8/basesWorkers8x8", "RANGED_RUSH", 0, 2, False),
    ("16x16/TwoBasesBarracks16x16", "HEAVY_DEFENSE", 1, 0, False),
    ("8x8/FourBasesWorkers8x8", "WORKER_DEFENSE", 0, 0, False),
    ("16x16/basesWorkers16x16", "WORKER_RUSH_PP", 1, 0, False),
    ("16x16/basesWorkers16x16", "PO_LIGHT_RUSH", 1, 0, True),
    ("8x8/basesWorkers8x8", "PO_WORKER_RUSH", 0, 1, True),
    ("16x16/basesWorkers16x16", "PO_RANGED_RUSH", 0, 2, True),
    ("24x24/basesWorkers24x24", "LIGHT_DEFENSE", 1, 0, True),
    ("16x16/basesWorkers16x16", "RANDOM_BIASED", 0, 0, True),
    ("16x16/basesWorkers16x16", "CRUSH_V1", 0, 0, False),
    ("8x8/basesWorkers8x8", "CRUSH_V1", 1, 0, False),
    ("24x24/basesWorkers24x24", "CRUSH_V1", 1, 0, True),
    ("16x16/basesWorkers16x16", "CRUSH_V2", 1, 0, False),
    ("16x16/TwoBasesBarracks16x16", "CRUSH_V2", 0, 1, True),
    ("16x16/basesWorkers16x16", "EMR_DETERMINISTICO", 0, 0, False),
    ("BWDistantResources32x32", "EMR_DETERMINISTICO", 1, 0, True),
]


@pytest.mark.parametrize("key,pol,side,pf,po", WIDE_SCRIPTED)
def test_wide_differential_scripted(backend, maps, key, pol, side, pf, po):
    import threading
    n = 3 if backend == "emu" else 2048
    total = 400 if backend == "emu" else 3000
    utt, outt = M.UnitTypeTable(1, 1), O.Utt(1, 1)
    b = M.BatchedGameState(utt, make_pgs(maps[key], utt), n, scripted_ai=True, po_policies=po)
    seeds = np.arange(n, dtype=np.int64) * 5 + 101 + SOAK
    b.reset(seeds)
    names = ["RANDOM_BIASED", "RANDOM_BIASED"]
    names[side] = pol
    kinds = [getattr(O, "AI_" + nm) for nm in names]
    for pl in range(2):
        b.set_policy(pl, getattr(M, "POLICY_" + names[pl]), pf)
    b.step(total, total)
    ex = b.export()
    games = [None] * n
    nthreads = min(16, os.cpu_count() or 1)

    def work(t):
        for g in range(t, n, nthreads):
            og = O.Game(outt, maps[key])
            og.seed(int(seeds[g]))
            ais = [O.ScriptedAI(k, pf) if k in O.SCRIPTED_AIS else None for k in kinds]
            (og.run_po if po else og.run)(kinds[0], ais[0], kinds[1], ais[1], total, total)
            games[g] = og

    ts = [threading.Thread(target=work, args=(t,)) for t in range(nthreads)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    for g, og in enumerate(games):
        P.assert_same_state(ex, g, og, "%s wide %s side %d pf %d po %s game %d" % (key, pol, side, pf, po, g))
    assert (b.results()[:, 3] == 0).all()
    b.close()


# ------------------------------------------------------------------------------------------------------------------
# PathFinding as an operator (mrts_batch_pathfind): the reference's test/microrts/TestPathfinding.java draws random
# destinations on random maps and compares two A* implementations; here the device's A*, BFS and greedy pathfinders answer
# random (unit, destination, range) queries on mid-game states and must return the oracle's first move
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("key", ["8x8/basesWorkers8x8", "16x16/basesWorkers16x16", "BWDistantResources32x32", "24x24/basesWorkers24x24", "GardenOfWar64x64"])
def test_pathfinding_operator(backend, maps, key):
    n = 4 if backend == "emu" else 256
    rounds = 2 if backend == "emu" else 6
    utt, outt = M.UnitTypeTable(1, 1), O.Utt(1, 1)
    b = M.BatchedGameState(utt, make_pgs(maps[key], utt), n, scripted_ai=True)
    seeds = np.arange(n, dtype=np.int64) + 900 + SOAK
    b.reset(seeds)
    b.set_policy(0, M.POLICY_RANDOM_BIASED)
    b.set_policy(1, M.POLICY_RANDOM_BIASED)
    games = []
    for g in range(n):
        og = O.Game(outt, maps[key])
        og.seed(int(seeds[g]))
        games.append(og)
    W, H = b.width, b.height
    rng = np.random.default_rng(7)
    for r in range(rounds):
        b.step(60, 3000)
        for og in games:
            og.run(O.AI_RANDOM_BIASED, None, O.AI_RANDOM_BIASED, None, 60, 3000)
        ex = b.export()
        idx = np.zeros(n, dtype=np.int64); cells = np.zeros(n, dtype=np.int32)
        for g in range(n):
            h, u, a = P.export_game(ex, g)
            idx[g] = rng.integers(0, len(u))
            cells[g] = u[idx[g], 2] + u[idx[g], 3] * W
        targets = rng.integers(0, W * H, size=n).astype(np.int32)
        # every other target next to some unit, so that short and blocked paths are common
        for g in range(0, n, 2):
            h, u, a = P.export_game(ex, g)
            j = rng.integers(0, len(u))
            targets[g] = int(u[j, 2]) + int(u[j, 3]) * W
        ranges = rng.choice(np.array([-1, 1, 1, 2, 3], dtype=np.int32), size=n)
        for pf in (M.PF_ASTAR, M.PF_BFS, M.PF_GREEDY):
            got = b.find_path(pf, cells, targets, ranges)
            for g, og in enumerate(games):
                want = og.pathfind(pf, int(idx[g]), int(targets[g]), int(ranges[g]))
                assert got[g] == want, "%s round %d game %d pf %d unit %d -> %d range %d: device %d oracle %d" % (
                    key, r, g, pf, idx[g], targets[g], ranges[g], got[g], want)
    b.close()


def test_evaluation_operator(backend, maps):
    """EvaluationFunction.evaluate as an operator (mrts_batch_evaluate): both evaluation functions, both players, fully
    observable and from each player's partially observable view, on mid-game states; float32 results must be identical."""
    key = "16x16/basesWorkers16x16"
    n = 4 if backend == "emu" else 128
    utt, outt = M.UnitTypeTable(1, 1), O.Utt(1, 1)
    b = M.BatchedGameState(utt, make_pgs(maps[key], utt), n)
    seeds = np.arange(n, dtype=np.int64) + 31 + SOAK
    b.reset(seeds)
    b.set_policy(0, M.POLICY_RANDOM_BIASED)
    b.set_policy(1, M.POLICY_RANDOM_BIASED)
    games = []
    for g in range(n):
        og = O.Game(outt, maps[key])
        og.seed(int(seeds[g]))
        games.append(og)
    for r in range(3):
        b.step(250, 3000)
        for og in games:
            og.run(O.AI_RANDOM_BIASED, None, O.AI_RANDOM_BIASED, None, 250, 3000)
        for fn in (0, 1):
            for maxp in (0, 1):
                for observer in (-1, 0, 1):
                    got = b.evaluate(fn, maxp, observer)
                    for g, og in enumerate(games):
                        view = og if observer < 0 else og.po_view(observer)
                        want = np.float32(view.evaluate(fn, maxp, 1 - maxp))
                        assert got[g] == want, "round %d fn %d maxplayer %d observer %d game %d: %r != %r" % (r, fn, maxp, observer, g, got[g], want)
    b.close()


# ------------------------------------------------------------------------------------------------------------------
# ordered unit action lists (Unit.getUnitActions) and the node loop of the MCTS searches, against the oracle
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("key,version", [("8x8/basesWorkers8x8", 1), ("16x16/basesWorkers16x16", 2), ("melee14x12Mixed18", 1), ("BWDistantResources32x32", 3)])
def test_unit_action_lists_and_cycle_to_decision(backend, maps, key, version):
    n = 3 if backend == "emu" else 16
    rounds = 5 if backend == "emu" else 14
    utt, outt = M.UnitTypeTable(version, 1), O.Utt(version, 1)
    m = maps[key]
    b = M.BatchedGameState(utt, make_pgs(m, utt), n)
    seeds = np.arange(n, dtype=np.int64) * 31 + 2 + SOAK
    b.reset(seeds)
    b.set_policy(0, M.POLICY_RANDOM_BIASED)
    b.set_policy(1, M.POLICY_RANDOM_BIASED)
    games = []
    for g in range(n):
        og = O.Game(outt, m)
        og.seed(int(seeds[g]))
        games.append(og)
    cost = [outt.field(t, 0) for t in range(7)]
    for r in range(rounds):
        for player in (0, 1):
            ua = b.unit_actions(player, none_duration=10 if r % 2 == 0 else 7)
            for g, og in enumerate(games):
                units, asg = og.units(), og.assignments()
                d = ua[g]
                assert d["time"] == og.time and d["gameover"] == og.gameover and d["winner"] == og.winner
                idle = [i for i in range(len(units)) if units[i][1] == player and not asg[i][0]]
                assert [c[0] for c in d["choices"]] == idle, (key, r, g, player)
                for (slot, uid, ty, x, y, acts) in d["choices"]:
                    assert (ty, x, y) == (units[slot][0], units[slot][2], units[slot][3]) and uid == units[slot][6]
                    ref = og.unit_actions(slot, 10 if r % 2 == 0 else 7)
                    norm = [(t, p, (ax if t == O.ATTACK else 0), (ay if t == O.ATTACK else 0), (ut if t == O.PRODUCE else -1)) for (t, p, ax, ay, ut) in ref]
                    got = [(t, (p if t != O.ATTACK else -1), ax, ay, ut) for (t, p, ax, ay, ut) in acts]
                    norm = [(t, (p if t != O.ATTACK else -1), ax, ay, ut) for (t, p, ax, ay, ut) in norm]
                    assert got == norm, "unit actions %s round %d game %d unit %d\ndev=%s\nref=%s" % (key, r, g, slot, got, norm)
                used, res_used = [], [0, 0]
                for i in range(len(units)):
                    if asg[i][0] and asg[i][1] in (O.MOVE, O.PRODUCE):
                        pos = units[i][2] + units[i][3] * m["w"] + {0: -m["w"], 1: 1, 2: m["w"], 3: -1}.get(int(asg[i][2]), 0)
                        used.append(int(pos))
                        if asg[i][1] == O.PRODUCE:
                            res_used[units[i][1]] += cost[asg[i][5]]
                assert d["positions_used"] == used and list(d["resources_used"]) == res_used
                assert d["resources"] == (og.resources(0), og.resources(1))
                assert d["can_act"] == tuple(any(units[i][1] == pl and not asg[i][0] for i in range(len(units))) for pl in (0, 1))
        # both players act, then the node loop: cycle until somebody can act again
        b.step(1, 3000)
        b.cycle_to_decision()
        ex = b.export()
        for g, og in enumerate(games):
            if not (og.gameover and og.time > 0):
                og.run(O.AI_RANDOM_BIASED, None, O.AI_RANDOM_BIASED, None, 1, 3000)
            while og.winner == -1 and not og.gameover and og.is_complete():
                og.cycle()
            P.assert_same_state(ex, g, og, "cycle_to_decision %s round %d" % (key, r))
    b.close()


@pytest.mark.parametrize("key,sizes,cycles,chunk", [("8x8/basesWorkers8x8", (6, 128), (3000, 3000), 53), ("8x8/FourBasesWorkers8x8", (4, 48), (1500, 3000), 7),
                                                     ("16x16/basesWorkers16x16", (3, 64), (900, 3000), 1), ("melee14x12Mixed18", (2, 32), (500, 2000), 10)])
def test_selfplay_through_the_observation_kernel(backend, maps, key, sizes, cycles, chunk):
    """Whole games through the fused step + observation kernel, whose shared-memory layout has no kind and no claim map
    (layout.h, slim): neighbour kinds come from grid[] + the unit table and the claims of cancelled pairs are coded into resv[].
    Small maps make same-cell conflicts between the two players frequent.  States, RNG streams and the final planes equal the oracle."""
    n = sizes[0] if backend == "emu" else sizes[1]
    total = cycles[0] if backend == "emu" else cycles[1]
    utt, outt = M.UnitTypeTable(1, 1), O.Utt(1, 1)
    m = maps[key]
    b = M.BatchedGameState(utt, make_pgs(m, utt), n)
    seeds = np.arange(n, dtype=np.int64) * 977 + 31 + SOAK
    b.reset(seeds)
    b.set_policy(0, M.POLICY_RANDOM_BIASED)
    b.set_policy(1, M.POLICY_RANDOM_BIASED)
    shape = (n, 6, m["h"], m["w"])
    o0, o1 = _device_buffer(backend, shape, np.uint8), _device_buffer(backend, shape, np.uint8)
    b.set_observation_outputs(o0, o1)
    games = []
    for g in range(n):
        og = O.Game(outt, m)
        og.seed(int(seeds[g]))
        games.append(og)
    every = max(1, 200 // chunk)
    for it, t in enumerate(range(0, total, chunk)):
        b.step(chunk, total)
        assert b.last_kernel == "k_step_fast_obs"
        for og in games:
            if not (og.gameover and og.time > 0):
                og.run(O.AI_RANDOM_BIASED, None, O.AI_RANDOM_BIASED, None, chunk, total)
        if it % every == 0 or t + chunk >= total:
            ex = b.export()
            a0, a1 = _to_numpy(o0), _to_numpy(o1)
            for g, og in enumerate(games):
                P.assert_same_state(ex, g, og, "%s obs kernel t=%d" % (key, t + chunk))
                assert [int(v) for v in ex["rng"][g]] == [og.rng_state(k) for k in range(3)]
                assert (a0[g] == og.observe(0).astype(np.uint8)).all() and (a1[g] == og.observe(1).astype(np.uint8)).all()
    b.close()
