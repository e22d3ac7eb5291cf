"""Parity of the CUDA engine (through the C ABI / microrts_b200) against the CPU oracle and the golden traces.

Every test here needs a device (`-m gpu`).  `MRTS_EMU=1` runs them against the emulated-warp debug build instead.
The bar is bit-exact: unit list (order, type, owner, position, resources, hit points, ids), player resources, time,
winner, and the in-flight assignments (action, issue time, insertion order).
"""
import numpy as np
import pytest

import microrts_b200 as M
import parity as P
from oracle import oracle as O

pytestmark = pytest.mark.gpu


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
    seeds = np.arange(n, dtype=np.int64) * 7919 + seed0
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
