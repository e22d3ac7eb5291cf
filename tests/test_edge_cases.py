"""Edge cases of the batched engine through the C ABI: empty and ragged inputs, capacity limits, odd batch sizes, finished
games, zero-length steps.  Each case is checked against the oracle where there is something to compare."""
import numpy as np
import pytest

import microrts_b200 as M
import parity as P
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def make(maps, key, n, **kw):
    utt = M.UnitTypeTable(1, 1)
    return utt, M.BatchedGameState(utt, M.PhysicalGameState.fromXML(P.map_to_xml(maps[key]), utt), n, **kw)


def test_empty_and_ragged_action_lists(backend, maps):
    """EXTERNAL players with no rows, zero counts, and counts that differ per game: games without actions only tick."""
    key = "8x8/basesWorkers8x8"
    utt, b = make(maps, key, 5)
    outt = O.Utt(1, 1)
    b.set_policy(0, M.POLICY_EXTERNAL)
    b.set_policy(1, M.POLICY_EXTERNAL)
    before = b.export()
    b.step(3, 3000)                                   # nothing staged at all
    after = b.export()
    assert (after["header"][:, 0] == 3).all() and (after["units"] == before["units"]).all()
    rows = np.zeros((5, 4, 8), dtype=np.int32)
    m = maps[key]
    games = [O.Game(outt, m) for _ in range(5)]
    for g in games:
        for _ in range(3):
            g.cycle()
    # worker of player 0 at its map position: harvest/move rows; counts 0,1,2,0,4 (rows past the count must be ignored)
    wx, wy = [(u[3], u[4]) for u in m["units"] if u[0] == "Worker" and u[2] == 0][0]
    cell = wx + wy * m["w"]
    for g in range(5):
        rows[g, 0] = [cell, 1, 2, 0, 0, 0, 0, 0]       # move down
        rows[g, 1] = [cell, 1, 1, 0, 0, 0, 0, 0]       # second row for the same unit: refused (unit no longer idle in pa)
        rows[g, 2] = [9999, 1, 0, 0, 0, 0, 0, 0]       # cell outside the map
        rows[g, 3] = [cell, 5, 0, 0, 0, 0, 0, 3]       # attack on an empty cell: illegal, issueSafe turns it into NONE
    counts = np.array([0, 1, 2, 0, 4], dtype=np.int32)
    b.set_actions(0, rows, counts, fill_none_duration=1)
    b.set_actions(1, np.zeros((5, 0, 8), dtype=np.int32).reshape(5, 0, 8), np.zeros(5, dtype=np.int32), fill_none_duration=1)
    b.step(1, 3000)
    ex = b.export()
    for g, og in enumerate(games):
        pa0 = og.from_vector_action(0, rows[g, :counts[g]], fill_none=1)
        pa1 = og.from_vector_action(1, np.zeros((0, 8), dtype=np.int32), fill_none=1)
        og.issue(pa0, True)
        og.issue(pa1, True)
        og.cycle()
        P.assert_same_state(ex, g, og, "ragged game %d" % g)
    assert (b.results()[:, 3] == 0).all()
    # a row with an action type outside 0..5 is dropped and flagged (MRTS_GE_BAD_ACTION = 16) instead of being interpreted
    bad = np.zeros((5, 1, 8), dtype=np.int32)
    bx, by = [(u[3], u[4]) for u in m["units"] if u[0] == "Base" and u[2] == 0][0]   # the base is idle again (its NONE(1) padding ran out)
    bad[:, 0] = [bx + by * m["w"], 7, 0, 0, 0, 0, 0, 0]
    b.set_actions(0, bad, np.array([0, 0, 1, 0, 0], dtype=np.int32), fill_none_duration=1)
    b.step(1, 3000)
    err = b.results()[:, 3]
    assert err[2] & 16 and (np.delete(err, 2) == 0).all()
    b.close()


def test_zero_cycles_and_finished_games(backend, maps):
    utt, b = make(maps, "8x8/basesWorkers8x8", 3)
    b.set_policy(0, M.POLICY_RANDOM_BIASED)
    b.set_policy(1, M.POLICY_RANDOM_BIASED)
    b.reset(np.array([1, 2, 3], dtype=np.int64))
    b.step(40, 3000)
    s0 = b.export()
    b.step(0, 3000)                                   # a zero-length step changes nothing
    s1 = b.export()
    for k in ("header", "units", "actions", "rng"):
        assert (s0[k] == s1[k]).all(), k
    b.step(3000, 300)                                 # the cap stops every game at time 300 at the latest
    r = b.results()
    assert (r[:, 0] <= 300).all() and ((r[:, 0] == 300) | (r[:, 2] == 1)).all()
    s2 = b.export()
    b.step(50, 300)                                   # finished / capped games are not advanced any further
    s3 = b.export()
    for k in ("header", "units", "actions", "rng"):
        assert (s2[k] == s3[k]).all(), k
    b.close()


@pytest.mark.parametrize("n", [1, 3, 37])
def test_odd_batch_sizes_match_any_position(backend, maps, n):
    """A game's trajectory depends on its seed only, not on the batch size or its position in the batch."""
    key = "16x16/basesWorkers16x16"
    utt, b = make(maps, key, n)
    utt2, ref = make(maps, key, 2)
    seeds = np.arange(n, dtype=np.int64) * 7 + 3
    for bb, s in ((b, seeds), (ref, seeds[[n - 1, 0]])):
        bb.set_policy(0, M.POLICY_RANDOM_BIASED)
        bb.set_policy(1, M.POLICY_RANDOM_BIASED)
        bb.reset(s)
        bb.step(250 if backend == "emu" else 900, 3000)
    a, r = b.export(), ref.export()
    for k in ("header", "units", "actions", "rng"):
        assert (a[k][n - 1] == r[k][0]).all() and (a[k][0] == r[k][1]).all(), k
    b.close(); ref.close()


def test_unit_capacity_overflow_is_flagged(backend, maps):
    """More live units than the batch's capacity: the produce is dropped and MRTS_GE_UNIT_OVERFLOW (bit 0) is set for that
    game, while games that stay within the capacity are unaffected (checked against the oracle)."""
    key = "8x8/basesWorkers8x8"
    utt = M.UnitTypeTable(1, 1)
    pgs = M.PhysicalGameState.fromXML(P.map_to_xml(maps[key]), utt)
    n_init = len(maps[key]["units"])
    with pytest.raises(M.MicroRTSError):
        M.BatchedGameState(utt, pgs, 2, unit_capacity=n_init - 1)   # fewer slots than initial units
    b = M.BatchedGameState(utt, pgs, 4, unit_capacity=32)
    assert b.cap == 32
    b.set_policy(0, M.POLICY_RANDOM_BIASED)
    b.set_policy(1, M.POLICY_RANDOM_BIASED)
    b.reset(np.arange(4, dtype=np.int64))
    b.step(3000, 3000)
    r = b.results()
    ex = b.export()
    assert (ex["header"][:, 3] <= 32).all()
    outt = O.Utt(1, 1)
    for g in range(4):
        og = O.Game(outt, maps[key]); og.seed(g)
        og.run(O.AI_RANDOM_BIASED, None, O.AI_RANDOM_BIASED, None, 3000, 3000)
        if r[g, 3] == 0:
            P.assert_same_state(ex, g, og, "within capacity")
        else:
            assert r[g, 3] & 1
    b.close()


def test_rollouts_ragged_roots_and_zero_depth(backend, maps):
    key = "8x8/basesWorkers8x8"
    utt, b = make(maps, key, 3)
    b.set_policy(0, M.POLICY_RANDOM_BIASED)
    b.set_policy(1, M.POLICY_RANDOM_BIASED)
    b.reset(np.array([5, 6, 7], dtype=np.int64))
    b.step(60, 3000)
    before = b.export()
    ev, tm = b.rollout(depth=0, rollouts_per_game=5)
    assert (tm == 0).all()
    outt = O.Utt(1, 1)
    for g in range(3):
        og = O.Game(outt, maps[key]); og.seed(5 + g)
        og.run(O.AI_RANDOM_BIASED, None, O.AI_RANDOM_BIASED, None, 60, 3000)
        assert (ev[g] == np.float32(og.evaluate(0, 0, 1))).all()
    after = b.export()
    for k in ("header", "units", "actions", "rng"):
        assert (before[k] == after[k]).all()
    b.close()


def test_duplicate_rows_cannot_overflow_the_pending_list(backend, maps):
    """max_k == unit capacity rows that all address the same idle unit (NONE actions are always accepted, as
    PlayerAction.fromVectorAction does): the pending list must not spill into the neighbouring shared-memory arrays.  The
    flooded games are flagged MRTS_GE_BAD_ACTION and stay well-formed; a game of the same batch with ordinary rows still equals
    the oracle."""
    key = "8x8/basesWorkers8x8"
    m = maps[key]
    utt, b = make(maps, key, 4)
    outt = O.Utt(1, 1)
    cap = b.cap
    b.set_policy(0, M.POLICY_EXTERNAL)
    b.set_policy(1, M.POLICY_EXTERNAL)
    cells = [[u[3] + u[4] * m["w"] for u in m["units"] if u[0] == "Worker" and u[2] == pl][0] for pl in (0, 1)]
    rows = [np.zeros((4, cap, 8), dtype=np.int32) for _ in range(2)]
    counts = [np.zeros(4, dtype=np.int32) for _ in range(2)]
    for pl in (0, 1):
        rows[pl][:, :, 0] = cells[pl]                  # every row: NONE for the player's worker
        counts[pl][:] = [cap, cap, 1, 3]               # games 0, 1 flooded; game 2 one row; game 3 three duplicates
    for pl in (0, 1):
        b.set_actions(pl, rows[pl], counts[pl], M.ACTIONS_VECTOR, fill_none_duration=1)
    b.step(1, 3000)
    ex = b.export()
    assert (ex["header"][:2, 6] & 16).all(), "flooded games are flagged"
    assert (ex["header"][2:, 6] == 0).all()
    for g in range(4):
        hdr, units, acts = P.export_game(ex, g)
        assert hdr[0] == 1 and hdr[3] == len(m["units"])
        assert not units[units[:, 1] < 0, 7].any(), "a neutral unit was given an assignment"
        exp = np.array([[P.TYPE_NAMES.index(u[0]), u[2], u[3], u[4], u[5], u[6]] for u in m["units"]], dtype=np.int32)
        assert (units[:, :6] == exp).all(), "unit table damaged"
    for g, k in ((2, 1), (3, 3)):
        og = O.Game(outt, m)
        pas = [og.from_vector_action(pl, rows[pl][g, :k], fill_none=1) for pl in (0, 1)]
        og.issue(pas[0], True); og.issue(pas[1], True); og.cycle()
        P.assert_same_state(ex, g, og, "duplicate rows game %d" % g)
    b.step(5, 3000)  # and the batch keeps stepping
    assert (b.export()["header"][:, 0] == 6).all()
    b.close()
