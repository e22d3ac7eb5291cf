"""The hand-derived answers of tests/test_oracle_golden.py (execute, move conflicts, evaluation, greedy pathfinding, observation and
mask of an initial state) asked of the DEVICE directly -- no oracle in the loop: the numbers below were worked out on paper from the
cited Java (UnitAction.java:307-465, GameState.java:249-328,922-968, UnitTypeTable.java, SimpleSqrtEvaluationFunction3.java:24-44,
GreedyPathFinding.java:53-84, UnitAction.java:711-751)."""
import numpy as np
import pytest

import microrts_b200 as M
import parity as P

pytestmark = pytest.mark.gpu

NONE, MOVE, HARVEST, RETURN, PRODUCE, ATTACK = range(6)


def tiny_map(units, w=8, h=8, walls=()):
    terrain = ["0"] * (w * h)
    for x, y in walls:
        terrain[x + y * w] = "1"
    return {"w": w, "h": h, "players": [[0, 5], [1, 5]], "terrain": "".join(terrain),
            "units": [[t, 100 + i, p, x, y, r, hp] for i, (t, p, x, y, r, hp) in enumerate(units)]}


def batch(units, version=1, conflict=1, scripted=False, **kw):
    utt = M.UnitTypeTable(version, conflict)
    b = M.BatchedGameState(utt, M.PhysicalGameState.fromXML(P.map_to_xml(tiny_map(units, **kw)), utt), 2, scripted_ai=scripted)
    b.reset(np.zeros(2, dtype=np.int64))
    return b


def raw(*rows):
    """one PlayerAction for both games of the batch: rows = (x, y, type, parameter, ax, ay, unit type)"""
    a = np.zeros((2, len(rows), 8), dtype=np.int32)
    for k, (x, y, t, p, ax, ay, ut) in enumerate(rows):
        a[:, k] = [x + y * 8, t, p, ax, ay, ut, 0, 0]
    return a


def state(b, g=0):
    ex = b.export()
    h, u, a = P.export_game(ex, g)
    return h, u, a


def test_execute(backend):
    b = batch([("Resource", -1, 0, 0, 20, 1), ("Worker", 0, 1, 0, 0, 1), ("Base", 0, 2, 0, 0, 10), ("Base", 1, 7, 7, 0, 10)])
    b.issue(0, raw((1, 0, HARVEST, 3, 0, 0, -1)))
    b.cycle(19)
    assert int(state(b)[1][1, 4]) == 0
    b.cycle(1)
    h, u, a = state(b)
    assert int(u[0, 4]) == 19 and int(u[1, 4]) == 1                  # harvest time 20, amount 1
    b.issue(0, raw((1, 0, RETURN, 1, 0, 0, -1)))
    b.cycle(10)                                                      # return takes the move time, 10
    h, u, a = state(b)
    assert [int(h[1]), int(h[2])] == [6, 5] and int(u[1, 4]) == 0
    b.issue(0, raw((2, 0, PRODUCE, 2, 0, 0, 3)))                     # the base trains a Worker below itself: 50 cycles
    b.cycle(49)
    h, u, a = state(b)
    assert [int(h[1]), int(h[2])] == [6, 5] and len(u) == 4
    b.cycle(1)
    h, u, a = state(b)
    assert [int(h[1]), int(h[2])] == [5, 5] and len(u) == 5 and u[4, :4].tolist() == [3, 0, 2, 1]
    b.close()
    # a Light (2 damage) kills the adjacent enemy Worker (1 hp) after attackTime 5; the worker's own assignment goes with it
    b = batch([("Light", 0, 1, 1, 0, 4), ("Worker", 1, 2, 1, 0, 1), ("Base", 0, 0, 7, 0, 10), ("Base", 1, 7, 7, 0, 10)])
    b.issue(0, raw((1, 1, ATTACK, -1, 2, 1, -1)))
    b.issue(1, raw((2, 1, MOVE, 1, 0, 0, -1)))
    b.cycle(5)
    h, u, a = state(b)
    assert len(u) == 3 and [int(t) for t in u[:, 0]] == [4, 1, 1] and u[:, 7].tolist() == [0, 0, 0]
    b.close()


@pytest.mark.parametrize("conflict", [1, 3])
def test_move_conflict_strategies(backend, conflict):
    units = [("Light", 0, 1, 1, 0, 4), ("Worker", 1, 3, 1, 0, 1), ("Base", 0, 0, 7, 0, 10), ("Base", 1, 7, 7, 0, 10),
             ("Light", 0, 1, 4, 0, 4), ("Worker", 1, 3, 4, 0, 1)]
    b = batch(units, conflict=conflict)
    b.issue(0, raw((1, 1, MOVE, 1, 0, 0, -1), (1, 4, MOVE, 1, 0, 0, -1)))   # both Lights step right (move time 8)
    b.issue(1, raw((3, 1, MOVE, 3, 0, 0, -1), (3, 4, MOVE, 3, 0, 0, -1)))   # both Workers step left (move time 10): same cells
    h, u, a = state(b)
    if conflict == 1:    # CANCEL_BOTH: every pair becomes NONE of duration min(8, 10)
        assert a[[0, 1, 4, 5], 0].tolist() == [NONE] * 4 and a[[0, 1, 4, 5], 1].tolist() == [8] * 4  # exported action = type, parameter, ...
    else:                # CANCEL_ALTERNATING: the first conflict cancels the new action, the second the old one
        assert a[[0, 1, 4, 5], 0].tolist() == [MOVE, NONE, NONE, MOVE]
    b.cycle(10)
    h, u, a = state(b)
    pos = [(int(r[2]), int(r[3])) for r in u]
    if conflict == 1:
        assert pos[0] == (1, 1) and pos[1] == (3, 1) and pos[4] == (1, 4) and pos[5] == (3, 4)
    else:
        assert pos[0] == (2, 1) and pos[1] == (3, 1) and pos[4] == (1, 4) and pos[5] == (2, 4)
    b.close()


def test_evaluation(backend):
    b = batch([("Base", 0, 1, 1, 0, 10), ("Worker", 0, 2, 2, 1, 1), ("Base", 1, 6, 6, 0, 4)])
    want = np.float32(2 * np.float32(550) / np.float32(650)) - np.float32(1)     # see test_oracle_golden.test_evaluation_known_answers
    assert b.evaluate(0, 0)[0] == np.float32(want)
    assert b.evaluate(1, 0)[0] == np.float32(290.0) and b.evaluate(1, 1)[0] == np.float32(-290.0)
    b.close()


def test_greedy_pathfinding(backend):
    units = [("Worker", 0, 1, 1, 0, 1), ("Base", 1, 6, 6, 0, 10)]
    b = batch(units, scripted=True)
    cells, tgt = [1 + 1 * 8] * 2, [5 + 1 * 8] * 2
    assert b.find_path(M.PF_GREEDY, cells, tgt, [1, 1]).tolist() == [1, 1]       # RIGHT is closest to (5,1)
    assert b.find_path(M.PF_GREEDY, cells, [3 + 1 * 8] * 2, [4, 1]).tolist() == [-1, 1]   # squared distance 4 <= "range" 4: null
    b.close()
    b = batch(units, scripted=True, walls=[(2, 1)])
    assert b.find_path(M.PF_GREEDY, cells, tgt, [1, 1]).tolist() == [0, 0]       # wall on the right: UP and DOWN tie, UP comes first
    b.close()
    b = batch(units, scripted=True, walls=[(2, 0), (2, 1), (2, 2)])
    assert b.find_path(M.PF_ASTAR, cells, [4 + 1 * 8] * 2, [-1, -1]).tolist() == [2, 2]   # around the wall's lower end
    assert b.find_path(M.PF_BFS, cells, [4 + 1 * 8] * 2, [-1, -1]).tolist() == [2, 2]
    b.close()


def test_initial_observation_and_masks(backend, maps):
    utt = M.UnitTypeTable(1, 1)
    b = M.BatchedGameState(utt, M.PhysicalGameState.fromXML(P.map_to_xml(maps["8x8/basesWorkers8x8"]), utt), 2)
    b.reset(np.zeros(2, dtype=np.int64))
    for player in (0, 1):
        want = np.zeros((6, 8, 8), dtype=np.int32)
        for (x, y, hp, res, owner, tid) in ((0, 0, 1, 20, -1, 0), (7, 7, 1, 20, -1, 0), (2, 1, 10, 0, 0, 1), (5, 6, 10, 0, 1, 1), (1, 1, 1, 0, 0, 3), (6, 6, 1, 0, 1, 3)):
            want[0, y, x], want[1, y, x], want[3, y, x] = hp, res, tid + 1
            if owner >= 0:
                want[2, y, x] = (owner + player) % 2 + 1
        assert (b.observe(player)[0] == want).all()
    m = b.masks(0)[0]
    w = np.zeros(79, dtype=np.int32)
    w[[0, 1 + NONE, 1 + MOVE, 1 + PRODUCE, 7 + 0, 7 + 2, 7 + 3, 19 + 0, 19 + 2, 19 + 3, 23 + 2]] = 1
    assert (m[1, 1] == w).all()
    bs = np.zeros(79, dtype=np.int32)
    bs[[0, 1 + NONE, 1 + PRODUCE, 19 + 0, 19 + 1, 19 + 2, 23 + 3]] = 1
    assert (m[1, 2] == bs).all()
    assert {(y, x) for y in range(8) for x in range(8) if m[y, x].any()} == {(1, 1), (1, 2)}
    b.close()


@pytest.mark.parametrize("pf,moves", [(0, (3, 2)), (1, (0, 1))])
def test_worker_rush_first_decision(backend, maps, pf, moves):
    """see test_oracle_golden.test_worker_rush_first_decision_known_answer: the bases train a Worker up / down, the workers step
    towards their resource -- LEFT / DOWN with A*, UP / RIGHT with BFS."""
    utt = M.UnitTypeTable(1, 1)
    b = M.BatchedGameState(utt, M.PhysicalGameState.fromXML(P.map_to_xml(maps["8x8/basesWorkers8x8"]), utt), 2, scripted_ai=True)
    b.reset(np.zeros(2, dtype=np.int64))
    b.set_policy(0, M.POLICY_WORKER_RUSH, pf)
    b.set_policy(1, M.POLICY_WORKER_RUSH, pf)
    b.step(1, 3000)
    h, u, a = state(b)
    assert a[2, [0, 1, 4]].tolist() == [PRODUCE, 0, 3] and a[3, [0, 1, 4]].tolist() == [PRODUCE, 2, 3]
    assert a[4, :2].tolist() == [MOVE, moves[0]] and a[5, :2].tolist() == [MOVE, moves[1]]
    b.close()


def test_partially_observable_observation(backend, maps):
    from test_oracle_golden import _po_expectations
    utt = M.UnitTypeTable(1, 1)
    b = M.BatchedGameState(utt, M.PhysicalGameState.fromXML(P.map_to_xml(maps["8x8/basesWorkers8x8"]), utt), 2, partial_obs=True)
    b.reset(np.zeros(2, dtype=np.int64))
    _po_expectations(b.observe(0)[0])
    b.close()
