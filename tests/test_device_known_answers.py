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


def _first_decision(units, policy, cycles=0, players=None, w=16, h=16):
    """the assignments one cycle after `policy` (player 0, A*) decided on the given state; player 1 is passive"""
    utt = M.UnitTypeTable(1, 1)
    mapd = tiny_map(units, w, h)
    if players is not None:
        mapd["players"] = players
    b = M.BatchedGameState(utt, M.PhysicalGameState.fromXML(P.map_to_xml(mapd), utt), 2, scripted_ai=True)
    b.reset(np.zeros(2, dtype=np.int64))
    b.set_policy(0, policy, M.PF_ASTAR)
    b.set_policy(1, M.POLICY_PASSIVE)
    if cycles:
        b.cycle(cycles)
    b.step(1, 3000)
    a = state(b)[2]
    b.close()
    return a


def test_crush_and_emr_decisions(backend):
    """The hand-derived answers of tests/test_oracle_golden.py (test_crush_v1_known_answers, test_crush_v2_known_answers,
    test_emr_deterministico_known_answers) asked of the device: cRush/CRush_V1.java + RangedAttack.java, CRush_V2.java +
    CRanged_Tactic.java, EMRDeterministico.java."""
    from test_oracle_golden import EMR_SECOND_BASE
    # CRush_V1 / RangedAttack: step back from a slower enemy well inside the range, shoot a faster one, shoot at the edge of the range
    racks = [("Base", 0, 0, 0, 0, 10), ("Barracks", 0, 0, 2, 0, 4), ("Ranged", 0, 8, 8, 0, 1), ("Heavy", 1, 8, 10, 0, 4), ("Base", 1, 15, 15, 0, 10)]
    a = _first_decision(racks, M.POLICY_CRUSH_V1)
    assert a[2, 0] == MOVE and a[2, 1] in (0, 3)
    a = _first_decision([u if u[0] != "Heavy" else ("Light", 1, 8, 10, 0, 4) for u in racks], M.POLICY_CRUSH_V1)
    assert a[2, 0] == ATTACK and a[2, 2:4].tolist() == [8, 10]
    a = _first_decision([u if u[0] != "Heavy" else ("Heavy", 1, 8, 11, 0, 4) for u in racks], M.POLICY_CRUSH_V1)
    assert a[2, 0] == ATTACK and a[2, 2:4].tolist() == [8, 11]
    # one running distance for "closest enemy" and "closest barracks": the Heavy listed after the barracks becomes the target
    shared = [("Base", 0, 0, 0, 0, 10), ("Ranged", 0, 8, 8, 0, 1), ("Heavy", 1, 8, 6, 0, 4), ("Barracks", 0, 0, 2, 0, 4), ("Heavy", 1, 8, 13, 0, 4), ("Base", 1, 15, 15, 0, 10)]
    a = _first_decision(shared, M.POLICY_CRUSH_V1)
    assert a[1, :2].tolist() == [MOVE, 2]
    # economy: 5 resources go to the barracks (base idle: NONE 10), 8 leave enough for "worker + ranged"
    eco = [("Base", 0, 2, 2, 0, 10), ("Worker", 0, 1, 2, 0, 1), ("Worker", 0, 1, 3, 0, 1), ("Worker", 0, 3, 3, 0, 1), ("Resource", -1, 0, 4, 20, 1),
           ("Base", 1, 13, 13, 0, 10), ("Worker", 1, 12, 12, 0, 1)]
    a = _first_decision(eco, M.POLICY_CRUSH_V1)
    assert a[0, :2].tolist() == [NONE, 10] and a[1, 0] == MOVE and a[1, 1] in (0, 3) and a[2, 0] == MOVE and a[3, 0] == MOVE and a[3, 1] in (1, 2)
    a = _first_decision(eco, M.POLICY_CRUSH_V1, players=[[0, 8], [1, 5]])
    assert a[0, [0, 4]].tolist() == [PRODUCE, 3]
    # CRush_V2 / CRanged_Tactic: the leader waits far from home, a follower lines up above the leader, stands still once there
    enemy = [("Base", 1, 15, 15, 0, 10), ("Worker", 1, 12, 8, 0, 1), ("Worker", 1, 14, 14, 0, 1)]
    a = _first_decision([("Base", 0, 0, 0, 0, 10), ("Ranged", 0, 8, 8, 0, 1)] + enemy, M.POLICY_CRUSH_V2)
    assert a[1, :2].tolist() == [NONE, 10]
    a = _first_decision([("Base", 0, 0, 0, 0, 10), ("Ranged", 0, 2, 2, 0, 1)] + enemy, M.POLICY_CRUSH_V2)
    assert a[1, 0] == MOVE and a[1, 1] in (1, 2)
    two = [("Base", 0, 0, 0, 0, 10), ("Ranged", 0, 8, 8, 0, 1), ("Ranged", 0, 4, 4, 0, 1)] + enemy
    a = _first_decision(two, M.POLICY_CRUSH_V2)
    assert a[2, 0] == MOVE and a[2, 1] in (1, 2)
    a = _first_decision([u if u[2:4] != (4, 4) else ("Ranged", 0, 8, 7, 0, 1) for u in two], M.POLICY_CRUSH_V2)
    assert a[2, :2].tolist() == [NONE, 10]
    a = _first_decision(two[:2] + [("Ranged", 0, 5, 8, 0, 1), ("Worker", 0, 8, 7, 0, 1)] + enemy, M.POLICY_CRUSH_V2)
    assert a[2, :2].tolist() == [MOVE, 1]
    # a Light attacks before cycle 400 and lines up behind the leading Ranged unit afterwards
    light = [("Base", 0, 0, 0, 0, 10), ("Ranged", 0, 8, 8, 0, 1), ("Light", 0, 8, 5, 0, 4)] + enemy
    a = _first_decision(light, M.POLICY_CRUSH_V2)
    assert a[2, 0] == MOVE and a[2, 1] in (1, 2)
    a = _first_decision(light, M.POLICY_CRUSH_V2, cycles=400)
    assert a[2, :2].tolist() == [MOVE, 2]
    # free workers leave a resource next to the enemy base alone
    a = _first_decision([("Base", 0, 0, 0, 0, 10), ("Worker", 0, 1, 1, 0, 1), ("Resource", -1, 1, 4, 20, 1)] + enemy, M.POLICY_CRUSH_V2)
    assert a[1, 0] == MOVE
    a = _first_decision([("Base", 0, 0, 0, 0, 10), ("Worker", 0, 1, 1, 0, 1), ("Resource", -1, 13, 13, 20, 1)] + enemy, M.POLICY_CRUSH_V2)
    assert a[1, 0] == NONE
    # EMRDeterministico: the second base goes next to the far resource that comes first in HashSet order (ID 112, listed last)
    a = _first_decision(EMR_SECOND_BASE, M.POLICY_EMR_DETERMINISTICO, players=[[0, 15], [1, 5]])
    assert a[2, :2].tolist() == [MOVE, 2] and a[0, [0, 4]].tolist() == [PRODUCE, 3] and a[1, [0, 4]].tolist() == [PRODUCE, 4]
    a = _first_decision(EMR_SECOND_BASE, M.POLICY_EMR_DETERMINISTICO, players=[[0, 7], [1, 5]])
    assert a[2, :2].tolist() == [MOVE, 1]
