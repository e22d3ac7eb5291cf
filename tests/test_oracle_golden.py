"""Pins the CPU oracle (oracle/mrts_oracle.c) against the reference's golden vectors.

* test_replay_all_traces: the protocol of test/microrts/TestTracesIntegrity.java:72-127, strengthened to assert
  state equality (units in list order + player resources) at every TraceEntry.
* test_regenerate_lightrush_traces: LightRush + AbstractionLayerAI + A* must reproduce the recorded actions of the
  140 LightRush mirror matches (src/tests/GenerateTestTraces.java:101-134).
* java.util.Random known answers (Java SE spec; values from SURVEY 8c).
"""
import numpy as np

from oracle import oracle as O


def test_java_random_known_answers():
    assert O.JavaRandom(42).next_int() == -1170105035
    assert O.JavaRandom(0).next_int() == -1155484576
    assert O.JavaRandom(42).next_double() == 0.7275636800328681
    r = O.JavaRandom(42)
    assert [r.next_int(10) for _ in range(8)] == [0, 3, 8, 4, 0, 5, 5, 8]


def test_utt_v1_matches_trace_utt(traces):
    utt = O.Utt(1, 1)
    fields, flags, prod = zip(*traces[0]["types"])
    for tid in range(7):
        assert [utt.field(tid, f) for f in range(12)] == list(fields[tid])
        assert utt.flags(tid) == flags[tid]
        assert utt.produces(tid) == list(prod[tid])


def replay(t, maps):
    utt = O.Utt.from_fields(t["conflict"], t["types"])
    g = O.Game(utt, maps[t["mapkey"]])
    for ei, e in enumerate(t["entries"]):
        while g.time < e["time"]:
            g.cycle()
        exp = np.array(e["units"], dtype=np.int32).reshape(-1, 6)
        got = g.units()[:, :6]
        assert got.shape == exp.shape and (got == exp).all(), (t["name"], ei, e["time"])
        assert (g.resources(0), g.resources(1)) == tuple(e["res"]), (t["name"], ei)
        if e["actions"]:
            p = [[], []]
            for (ui, ty, par, x, y, ut) in e["actions"]:
                p[exp[ui][1]].append((ui, (ty, par, x, y, ut)))
            g.issue(p[0], True)
            g.issue(p[1], True)
    assert g.errors == 0
    return g


def test_replay_all_traces(traces, maps):
    assert len(traces) == 280
    for t in traces:
        replay(t, maps)


def _norm(a):
    ty, par, x, y, ut = a
    return (ty, -1, x, y, -1) if ty == O.ATTACK else (ty, par, 0, 0, ut)


def test_regenerate_lightrush_traces(traces, maps):
    n = 0
    for t in traces:
        if "LightRush" not in t["name"]:
            continue
        n += 1
        utt = O.Utt.from_fields(t["conflict"], t["types"])
        g = O.Game(utt, maps[t["mapkey"]])
        ais = [O.ScriptedAI(O.AI_LIGHT_RUSH), O.ScriptedAI(O.AI_LIGHT_RUSH)]
        ents, ei, gameover = t["entries"], 1, False
        while not gameover and g.time < 500:
            pa0, pa1 = ais[0].get_action(g, 0), ais[1].get_action(g, 1)
            if pa0 or pa1:
                e = ents[ei]
                ei += 1
                assert e["time"] == g.time, t["name"]
                exp = [(ui, _norm((ty, par, x, y, ut))) for (ui, ty, par, x, y, ut) in e["actions"]]
                assert [(u, _norm(a)) for u, a in pa0 + pa1] == exp, (t["name"], g.time)
            g.issue(pa0, True)
            g.issue(pa1, True)
            gameover = g.cycle()
        assert ei == len(ents) - 1, t["name"]
        last = np.array(ents[-1]["units"], dtype=np.int32).reshape(-1, 6)
        assert (g.units()[:, :6] == last).all()
    assert n == 140


# ------------------------------------------------------------------------------------------------------------------
# hand-derived known answers for the parts of the oracle the reference holds no golden data for (worked out on paper from
# the cited Java: GreedyPathFinding.java:53-84, PartiallyObservableGameState.java:35-71, WorkerDefense.java:117-146,
# POLightRush.java:42-78).  They keep the checker itself honest on CPU; the device is compared with it in the gpu tests.
# ------------------------------------------------------------------------------------------------------------------
def _tiny_map(units, w=8, h=8, walls=()):
    terrain = ["0"] * (w * h)
    for x, y in walls:
        terrain[x + y * w] = "1"
    return {"w": w, "h": h, "players": [[0, 5], [1, 5]], "terrain": "".join(terrain),
            "units": [[t, 100 + i, p, x, y, r, hp] for i, (t, p, x, y, r, hp) in enumerate(units)]}


def test_greedy_pathfinder_known_answers():
    utt = O.Utt(1, 1)
    # a worker at (1,1), target (5,1): RIGHT is the free neighbour closest to the target
    g = O.Game(utt, _tiny_map([("Worker", 0, 1, 1, 0, 1), ("Base", 1, 6, 6, 0, 10)]))
    assert g.pathfind(O.PF_GREEDY, 0, 5 + 1 * 8, 1) == 1
    # a wall on the right: UP (1,0) and DOWN (1,2) tie at squared distance 17, LEFT is farther; the first in direction order wins
    g = O.Game(utt, _tiny_map([("Worker", 0, 1, 1, 0, 1), ("Base", 1, 6, 6, 0, 10)], walls=[(2, 1)]))
    assert g.pathfind(O.PF_GREEDY, 0, 5 + 1 * 8, 1) == 0
    # "already in range" compares the SQUARED distance with the unsquared range (GreedyPathFinding.java:66): at squared distance 4
    # a range of 3 returns null although 2 <= 3 would also, and a range of 1 does not (distance 2 > 1)
    g = O.Game(utt, _tiny_map([("Worker", 0, 1, 1, 0, 1), ("Base", 1, 6, 6, 0, 10)]))
    assert g.pathfind(O.PF_GREEDY, 0, 3 + 1 * 8, 4) == -1
    assert g.pathfind(O.PF_GREEDY, 0, 3 + 1 * 8, 1) == 1
    # boxed in by units and the map edge: null
    g = O.Game(utt, _tiny_map([("Worker", 0, 0, 0, 0, 1), ("Worker", 0, 1, 0, 0, 1), ("Worker", 0, 0, 1, 0, 1), ("Base", 1, 6, 6, 0, 10)]))
    assert g.pathfind(O.PF_GREEDY, 0, 5 + 5 * 8, 1) == -1
    # A* and BFS agree with the obvious first step around a wall segment
    g = O.Game(utt, _tiny_map([("Worker", 0, 1, 1, 0, 1), ("Base", 1, 6, 6, 0, 10)], walls=[(2, 0), (2, 1), (2, 2)]))
    assert g.pathfind(O.PF_ASTAR, 0, 4 + 1 * 8, 0) == g.pathfind(O.PF_BFS, 0, 4 + 1 * 8, 0) == 2  # down, around the wall's lower end


def test_floodfill_pathfinder_known_answers():
    """FloodFillPathFinding.java:47-107,139-213 by hand on an empty 8x8 map."""
    utt = O.Utt(1, 1)
    ff = O.FloodFill()
    target = 5 + 1 * 8
    # worker at (1,1), target (5,1): the flood from the target reaches (2,1) at distance 3 and stops after expanding it (its
    # neighbourhood holds the start), so (1,0) and (1,2) (distance 5) are never labelled: RIGHT
    g = O.Game(utt, _tiny_map([("Worker", 0, 1, 1, 0, 1), ("Base", 1, 6, 6, 0, 10)]))
    assert ff.find(g, 0, target, 1) == 1
    # already within range 1 of the target: null, and nothing is cached
    g2 = O.Game(utt, _tiny_map([("Worker", 0, 4, 1, 0, 1), ("Base", 1, 6, 6, 0, 10)]))
    assert ff.find(g2, 0, target, 1) == -1
    # same instance, same target, but (2,1) is now occupied: the cached map still says RIGHT, the cell is not free, so the map is
    # dropped and recomputed around the blocker; (1,0) and (1,2) are both labelled 5 before either is expanded -> first minimum in
    # the order left, up, right, down = UP
    g3 = O.Game(utt, _tiny_map([("Worker", 0, 1, 1, 0, 1), ("Worker", 0, 2, 1, 0, 1), ("Base", 1, 6, 6, 0, 10)]))
    assert ff.find(g3, 0, target, 1) == 0
    # a fresh instance with (1,0) reserved by an earlier desire of the same cycle (ru): DOWN
    assert O.FloodFill().find(g3, 0, target, 1, ru=[1 + 0 * 8]) == 2
    # the stale map is reused as long as the step it suggests is free: the instance that computed "UP" for g3 answers UP for the
    # unblocked map g as well (a fresh instance says RIGHT)
    assert ff.find(g, 0, target, 1) == 0 and O.FloodFill().find(g, 0, target, 1) == 1


def test_partially_observable_view_known_answers():
    utt = O.Utt(1, 1)
    # worker sight radius 3, base 5 (UnitTypeTable.java): player 0 has a worker at (1,1) only
    units = [("Worker", 0, 1, 1, 0, 1), ("Resource", -1, 4, 1, 20, 1), ("Resource", -1, 5, 1, 20, 1), ("Worker", 1, 3, 3, 0, 1), ("Base", 1, 6, 6, 0, 10)]
    g = O.Game(utt, _tiny_map(units))
    v = g.po_view(0)
    seen = {(int(u[2]), int(u[3])) for u in v.units()}
    # (4,1): dx=3 -> 9 <= 9 visible; (5,1): 16 > 9 hidden; (3,3): 4+4=8 visible; (6,6): hidden; own unit always there
    assert seen == {(1, 1), (4, 1), (3, 3)}
    v1 = g.po_view(1)
    seen1 = {(int(u[2]), int(u[3])) for u in v1.units()}
    # player 1: worker (3,3) radius 3 and base (6,6) radius 5: (1,1): 8 <= 9 from the worker; (4,1): 1+4=5; (5,1): 4+4=8
    assert seen1 == {(1, 1), (4, 1), (5, 1), (3, 3), (6, 6)}


def _action_of(pairs, unit_idx):
    for u, a in pairs:
        if u == unit_idx:
            return a
    return None


def test_defense_and_po_rush_known_answers():
    utt = O.Utt(1, 1)
    # 8x8 map (height / 2 = 4).  Player 0: base (0,0), a Light at (1,0) [distance to base 1 < 4].  Enemy worker far away at (7,7).
    near_base = [("Base", 0, 0, 0, 0, 10), ("Light", 0, 1, 0, 0, 4), ("Worker", 1, 7, 7, 0, 1), ("Base", 1, 6, 7, 0, 10)]
    g = O.Game(utt, _tiny_map(near_base))
    ai = O.ScriptedAI(O.AI_LIGHT_DEFENSE)
    act = _action_of(ai.get_action(g, 0), 1)
    assert act is not None and act[0] == O.MOVE            # close to its base: it goes after the closest enemy (LightDefense.java:157-159)
    # the same Light far from its base (6,0) [distance 6] and far from every enemy (closest: (7,7), distance 8): Attack(null),
    # deleted by translateActions, so fillWithNones gives it NONE(10) (LightDefense.java:160-163, PlayerAction.java:217-235)
    far = [("Base", 0, 0, 0, 0, 10), ("Light", 0, 6, 0, 0, 4), ("Worker", 1, 7, 7, 0, 1), ("Base", 1, 6, 7, 0, 10)]
    g = O.Game(utt, _tiny_map(far))
    ai = O.ScriptedAI(O.AI_LIGHT_DEFENSE)
    act = _action_of(ai.get_action(g, 0), 1)
    assert act is not None and act[0] == O.NONE and act[1] == 10
    # WorkerRushPlusPlus attacks regardless of the distances (WorkerRushPlusPlus.java:136-138)
    ai = O.ScriptedAI(O.AI_WORKER_RUSH_PP)
    act = _action_of(ai.get_action(O.Game(utt, _tiny_map(far)), 0), 1)
    assert act is not None and act[0] == O.MOVE
    # POLightRush on player 0's view of a 16x16 map: no enemy in sight, so the Light at (2,2) explores towards the nearest cell
    # nobody sees.  Sight: base (0,0) radius 5, Light radius 2 -> the first non-observable cell in row-major order with the
    # smallest squared distance to (2,2) is (4,4)?  (dx,dy)=(2,2): 8 > 4 for the Light; from the base 16+16=32 > 25 -> unseen,
    # distance 8; closer candidates: (2,5)/(5,2): distance 9; (3,4): base 9+16=25 seen; (4,3) seen -> target (4,4): move DOWN
    # or RIGHT first; A* expands ... the move must at least be a MOVE towards larger x or y.
    units = [("Base", 0, 0, 0, 0, 10), ("Light", 0, 2, 2, 0, 4), ("Base", 1, 15, 15, 0, 10)]
    g = O.Game(utt, _tiny_map(units, 16, 16))
    ai = O.ScriptedAI(O.AI_PO_LIGHT_RUSH)
    act = _action_of(ai.get_action(g.po_view(0), 0), 1)
    assert act is not None and act[0] == O.MOVE and act[1] in (1, 2)
    # on the fully observable state the same class attacks instead (POLightRush.java:52-54)
    ai = O.ScriptedAI(O.AI_PO_LIGHT_RUSH)
    act = _action_of(ai.get_action(g, 0), 1)
    assert act is not None and act[0] == O.MOVE


def test_crush_v1_known_answers():
    """cRush/CRush_V1.java + RangedAttack.java, decided by hand on tiny 16x16 states (more than 144 cells: the barracks
    variant).  Type table version 1: Ranged range 3 / move time 10, Heavy move time 12, Light move time 8."""
    utt = O.Utt(1, 1)
    # -- RangedAttack.execute (:58-87).  Ranged at (8,8), own barracks at (0,2): sqrt(64+36) > 2.
    # a Heavy two cells below: d = 2 <= range - 1 and 10 < 12 -> it walks towards the barracks (up or left), away from the Heavy
    racks = [("Base", 0, 0, 0, 0, 10), ("Barracks", 0, 0, 2, 0, 4), ("Ranged", 0, 8, 8, 0, 1), ("Heavy", 1, 8, 10, 0, 4), ("Base", 1, 15, 15, 0, 10)]
    act = _action_of(O.ScriptedAI(O.AI_CRUSH_V1).get_action(O.Game(utt, _tiny_map(racks, 16, 16)), 0), 2)
    assert act is not None and act[0] == O.MOVE and act[1] in (0, 3)
    # a Light in the same place is not slower (8 < 10): d <= range -> shoot it
    light = [u if u[0] != "Heavy" else ("Light", 1, 8, 10, 0, 4) for u in racks]
    act = _action_of(O.ScriptedAI(O.AI_CRUSH_V1).get_action(O.Game(utt, _tiny_map(light, 16, 16)), 0), 2)
    assert act is not None and act[0] == O.ATTACK and (act[2], act[3]) == (8, 10)
    # the Heavy at distance 3 = range: not "well inside" (3 > range - 1) -> shoot
    edge = [u if u[0] != "Heavy" else ("Heavy", 1, 8, 11, 0, 4) for u in racks]
    act = _action_of(O.ScriptedAI(O.AI_CRUSH_V1).get_action(O.Game(utt, _tiny_map(edge, 16, 16)), 0), 2)
    assert act is not None and act[0] == O.ATTACK and (act[2], act[3]) == (8, 11)
    # without a barracks (racks == null, rd = 0) it shoots the adjacent Heavy as well
    none = [u for u in racks if u[0] != "Barracks"]
    act = _action_of(O.ScriptedAI(O.AI_CRUSH_V1).get_action(O.Game(utt, _tiny_map(none, 16, 16)), 0), 1)
    assert act is not None and act[0] == O.ATTACK
    # -- rangedUnitBehavior (:194-220) shares ONE running distance between "closest enemy" and "closest barracks": Heavy A at
    # distance 2 comes first, then the barracks (always accepted as the first one) sets the distance to 14, and Heavy B at
    # distance 5, listed later, is now "closer" and becomes the target: 5 > range -> the unit walks DOWN towards B instead of
    # stepping back from A
    shared = [("Base", 0, 0, 0, 0, 10), ("Ranged", 0, 8, 8, 0, 1), ("Heavy", 1, 8, 6, 0, 4), ("Barracks", 0, 0, 2, 0, 4), ("Heavy", 1, 8, 13, 0, 4), ("Base", 1, 15, 15, 0, 10)]
    act = _action_of(O.ScriptedAI(O.AI_CRUSH_V1).get_action(O.Game(utt, _tiny_map(shared, 16, 16)), 0), 1)
    assert act is not None and act[0] == O.MOVE and act[1] == 2
    # -- workersBehavior / baseBehavior (:222-321, :133-168): base (2,2), three workers, 5 resources, no barracks.  nbases + 1 = 2
    # workers stay free: the first is sent to build the barracks (nworkers > 1, 5 >= 5) at the first free cell of the ring around
    # it, (0,1): a diagonal neighbour of (1,2), so it moves up or left; the second harvests from (0,4): it walks (down or left);
    # the third is a battle worker and walks towards the enemy.  The base: 3 workers >= nbases + 1; resourcesUsed = 5 != 5 * 0
    # barracks -> 5 - 5 = 0 resources left for "worker + ranged" -> nothing to do: NONE(10)
    eco = [("Base", 0, 2, 2, 0, 10), ("Worker", 0, 1, 2, 0, 1), ("Worker", 0, 1, 3, 0, 1), ("Worker", 0, 3, 3, 0, 1), ("Resource", -1, 0, 4, 20, 1),
           ("Base", 1, 13, 13, 0, 10), ("Worker", 1, 12, 12, 0, 1)]
    m = _tiny_map(eco, 16, 16)
    pa = O.ScriptedAI(O.AI_CRUSH_V1).get_action(O.Game(utt, m), 0)
    assert _action_of(pa, 0)[0] == O.NONE and _action_of(pa, 0)[1] == 10
    assert _action_of(pa, 1)[0] == O.MOVE and _action_of(pa, 1)[1] in (0, 3)
    assert _action_of(pa, 2)[0] == O.MOVE and _action_of(pa, 2)[1] in (2, 3)
    assert _action_of(pa, 3)[0] == O.MOVE and _action_of(pa, 3)[1] in (1, 2)
    # with 8 resources the base sees 8 - 5 = 3 >= worker (1) + ranged (2) and, buildingRacks being set, trains a worker
    m8 = dict(m, players=[[0, 8], [1, 5]])
    pa = O.ScriptedAI(O.AI_CRUSH_V1).get_action(O.Game(utt, m8), 0)
    assert _action_of(pa, 0)[0] == O.PRODUCE and _action_of(pa, 0)[4] == 3
    # -- on a map of at most 144 cells the same state is a worker rush (:329-416): one harvester per base, the other two fight
    # (no barracks is ever planned), and the base trains workers whenever it can pay for one
    pa = O.ScriptedAI(O.AI_CRUSH_V1).get_action(O.Game(utt, _tiny_map([u if u[1] != 1 else (u[0], 1, u[2] - 4, u[3] - 4, u[4], u[5]) for u in eco], 12, 12)), 0)
    assert _action_of(pa, 0)[0] == O.PRODUCE and _action_of(pa, 0)[4] == 3
    assert _action_of(pa, 1)[0] == O.MOVE and _action_of(pa, 1)[1] in (2, 3)      # the harvester (first worker) heads for (0,4)
    assert _action_of(pa, 2)[0] == O.MOVE and _action_of(pa, 2)[1] in (1, 2)      # battle workers walk towards the enemy
    assert _action_of(pa, 3)[0] == O.MOVE and _action_of(pa, 3)[1] in (1, 2)


def test_crush_v2_known_answers():
    """cRush/CRush_V2.java + CRanged_Tactic.java, decided by hand on 16x16 states (Ranged: range 3, sight 3)."""
    utt = O.Utt(1, 1)

    def decide(units, unit_idx, cycles=0, player=0):
        g = O.Game(utt, _tiny_map(units, 16, 16))
        for _ in range(cycles):
            g.cycle()
        return _action_of(O.ScriptedAI(O.AI_CRUSH_V2).get_action(g, player), unit_idx)

    # enemy: a base and two workers -> not "time to attack" (2 workers >= 2 * 1 bases)
    enemy = [("Base", 1, 15, 15, 0, 10), ("Worker", 1, 12, 8, 0, 1), ("Worker", 1, 14, 14, 0, 1)]
    # the only Ranged unit leads.  Far from home (distance^2 128 >= 25) and closer to the enemy base (98) than home is (450): it
    # waits (move == null -> NONE(10), :186-193) ...
    act = decide([("Base", 0, 0, 0, 0, 10), ("Ranged", 0, 8, 8, 0, 1)] + enemy, 1)
    assert act[0] == O.NONE and act[1] == 10
    # ... close to home (8 < 25) it heads for the enemy base
    act = decide([("Base", 0, 0, 0, 0, 10), ("Ranged", 0, 2, 2, 0, 1)] + enemy, 1)
    assert act[0] == O.MOVE and act[1] in (1, 2)
    # a target within range is shot whatever the role (:179-181)
    act = decide([("Base", 0, 0, 0, 0, 10), ("Ranged", 0, 10, 8, 0, 1)] + enemy, 1)
    assert act[0] == O.ATTACK and (act[2], act[3]) == (12, 8)
    # a second Ranged unit follows the one nearest to the enemy base, (8,8).  The cell below the leader is closer to the enemy base
    # (85 < 98), so the formation grows up / left: first the cell above the leader, (8,7) -- the follower at (4,4) walks right or down
    two = [("Base", 0, 0, 0, 0, 10), ("Ranged", 0, 8, 8, 0, 1), ("Ranged", 0, 4, 4, 0, 1)] + enemy
    act = decide(two, 2)
    assert act[0] == O.MOVE and act[1] in (1, 2)
    # standing on (8,7) it is in formation: null -> NONE(10) (:326-328)
    act = decide([u if u[2:4] != (4, 4) else ("Ranged", 0, 8, 7, 0, 1) for u in two], 2)
    assert act[0] == O.NONE and act[1] == 10
    # with (8,7) taken by another unit the next free cell is the one on the left, (7,8): the follower at (5,8) walks right
    act = decide(two[:2] + [("Ranged", 0, 5, 8, 0, 1), ("Worker", 0, 8, 7, 0, 1)] + enemy, 2)
    assert act[0] == O.MOVE and act[1] == 1
    # one enemy worker left for one base and no combat unit: time to attack, the follower goes after its target (the worker at (12,8))
    act = decide([("Base", 0, 0, 0, 0, 10), ("Ranged", 0, 8, 8, 0, 1), ("Ranged", 0, 4, 8, 0, 1), ("Base", 1, 15, 15, 0, 10), ("Worker", 1, 12, 8, 0, 1)], 2)
    assert act[0] == O.MOVE and act[1] == 1
    # before cycle 400 a Light unit simply attacks (:213-214); from cycle 400 on it follows the tactic like a Ranged unit -- the
    # leader is the Ranged unit (8,8), and the Light at (8,5) is already where a follower belongs?  No: (8,7) is free, so it walks
    # down towards it
    light = [("Base", 0, 0, 0, 0, 10), ("Ranged", 0, 8, 8, 0, 1), ("Light", 0, 8, 5, 0, 4)] + enemy
    act = decide(light, 2)
    assert act[0] == O.MOVE and act[1] in (1, 2)           # Attack: A* towards the worker at (12,8)
    act = decide(light, 2, cycles=400)
    assert act[0] == O.MOVE and act[1] == 2                # formation: down to (8,7)
    # free workers (:335-383): the closest resource (1,4) lies nearer to the own base (0,0) than to the enemy base -> harvest;
    # a resource next to the enemy base instead is left alone, and without an enemy base nobody harvests at all
    eco = [("Base", 0, 0, 0, 0, 10), ("Worker", 0, 1, 1, 0, 1), ("Resource", -1, 1, 4, 20, 1)] + enemy
    assert decide(eco, 1)[0] == O.MOVE
    far = [("Base", 0, 0, 0, 0, 10), ("Worker", 0, 1, 1, 0, 1), ("Resource", -1, 13, 13, 20, 1)] + enemy
    assert decide(far, 1)[0] == O.NONE
    nobase = [("Base", 0, 0, 0, 0, 10), ("Worker", 0, 1, 1, 0, 1), ("Resource", -1, 1, 4, 20, 1), ("Worker", 1, 12, 8, 0, 1)]
    assert decide(nobase, 1)[0] == O.NONE


EMR_SECOND_BASE = [("Base", 0, 1, 1, 0, 10), ("Barracks", 0, 3, 1, 0, 4), ("Worker", 0, 10, 8, 0, 1), ("Resource", -1, 0, 0, 20, 1),
                   ("Resource", -1, 13, 8, 20, 1),                                            # ID 104: hash slot 104 & 15 = 8
                   ("Base", 1, 15, 15, 0, 10), ("Worker", 1, 14, 15, 0, 1), ("Worker", 1, 13, 15, 0, 1), ("Worker", 1, 12, 15, 0, 1),
                   ("Worker", 1, 11, 15, 0, 1), ("Worker", 1, 10, 15, 0, 1), ("Worker", 1, 9, 15, 0, 1),
                   ("Resource", -1, 9, 13, 20, 1)]                                            # ID 112: hash slot 0


def test_emr_deterministico_known_answers():
    """EMRDeterministico.java: the second base goes next to the FIRST ELEMENT OF A HashSet<Unit> of far resources (:264-276,
    :287-311).  Unit.hashCode() is the ID, so with IDs 104 and 112 in a 16-slot table the resource listed LAST (slot 0) comes
    first: desired position (10,14); the ring of radius 1 around it starts with its top row y = 13, x = 9..11, where (9,13) holds
    the resource and (10,13) is free.  The worker at (10,8) walks straight down; for the other resource, (13,8), it would walk
    right."""
    utt = O.Utt(1, 1)
    m = dict(_tiny_map(EMR_SECOND_BASE, 16, 16), players=[[0, 15], [1, 5]])
    pa = O.ScriptedAI(O.AI_EMR_DETERMINISTICO).get_action(O.Game(utt, m), 0)
    act = _action_of(pa, 2)
    assert act[0] == O.MOVE and act[1] == 2
    # the base trains a worker (1 < 6 workers), the barracks a Light (0 army units: turn 0)
    assert _action_of(pa, 0)[0] == O.PRODUCE and _action_of(pa, 0)[4] == 3
    assert _action_of(pa, 1)[0] == O.PRODUCE and _action_of(pa, 1)[4] == 4
    # with too few resources for a base (10) nothing is built and the worker harvests from the closest resource, (13,8): right
    poor = dict(m, players=[[0, 7], [1, 5]])
    act = _action_of(O.ScriptedAI(O.AI_EMR_DETERMINISTICO).get_action(O.Game(utt, poor), 0), 2)
    assert act[0] == O.MOVE and act[1] == 1
    # play it: the new base appears in the lower part of the map, next to resource 112, not next to 104
    g = O.Game(utt, m)
    ai = O.ScriptedAI(O.AI_EMR_DETERMINISTICO)
    g.run(O.AI_EMR_DETERMINISTICO, ai, O.AI_PASSIVE, None, 400, 3000)
    bases = sorted((int(u[2]), int(u[3])) for u in g.units() if u[0] == 1 and u[1] == 0)
    assert (1, 1) in bases and len(bases) >= 2 and all(b == (1, 1) or b[1] >= 12 for b in bases), bases


def test_observation_and_mask_known_answers(maps):
    """GameState.getVectorObservation (GameState.java:922-968) and JNIGridnetClient.getMasks / UnitAction.getValidActionArray
    (UnitAction.java:711-751) of the initial state of maps/8x8/basesWorkers8x8.xml, written out by hand."""
    utt = O.Utt(1, 1)
    g = O.Game(utt, maps["8x8/basesWorkers8x8"])
    # units: Resource (0,0) and (7,7) with 20 resources and 1 hp; Base p0 (2,1) hp 10; Base p1 (5,6); Worker p0 (1,1); Worker p1 (6,6)
    for player in (0, 1):
        o = g.observe(player)
        want = np.zeros((6, 8, 8), dtype=np.int32)
        for (x, y, hp, res, owner, tid) in ((0, 0, 1, 20, -1, 0), (7, 7, 1, 20, -1, 0), (2, 1, 10, 0, 0, 1), (5, 6, 10, 0, 1, 1), (1, 1, 1, 0, 0, 3), (6, 6, 1, 0, 1, 3)):
            want[0, y, x] = hp
            want[1, y, x] = res
            if owner >= 0:
                want[2, y, x] = (owner + player) % 2 + 1  # 1 = the observing player's own unit, 2 = the opponent's
            want[3, y, x] = tid + 1
        assert (o == want).all()                          # no action in flight (plane 4), no walls (plane 5)
    m = g.masks(0)
    assert m.shape == (8, 8, 1 + 6 + 16 + 7 + 49)
    # worker (1,1) of player 0, 5 resources: up (1,0), down (1,2) and left (0,1) are free, the base is on the right; the resource at
    # (0,0) is diagonal (no harvest); it can afford a Barracks (5) but not a Base (10)
    w = np.zeros(79, dtype=np.int32)
    w[0] = 1
    w[1 + O.NONE] = w[1 + O.MOVE] = w[1 + O.PRODUCE] = 1
    w[7 + 0] = w[7 + 2] = w[7 + 3] = 1                    # move directions
    w[19 + 0] = w[19 + 2] = w[19 + 3] = 1                 # produce directions
    w[23 + 2] = 1                                         # produce type: Barracks (ID 2)
    assert (m[1, 1] == w).all()
    # base (2,1): trains a Worker (cost 1) into (2,0), (3,1) or (2,2); the worker stands on its left
    b = np.zeros(79, dtype=np.int32)
    b[0] = 1
    b[1 + O.NONE] = b[1 + O.PRODUCE] = 1
    b[19 + 0] = b[19 + 1] = b[19 + 2] = 1
    b[23 + 3] = 1
    assert (m[1, 2] == b).all()
    mask_cells = {(y, x) for y in range(8) for x in range(8) if m[y, x].any()}
    assert mask_cells == {(1, 1), (1, 2)}                 # only player 0's idle units carry a mask


def test_evaluation_known_answers():
    """SimpleSqrtEvaluationFunction3.java:24-44 (hp / maxhp is an INTEGER division inside the sqrt) and
    SimpleEvaluationFunction.java:21-36, on a position small enough to add up by hand."""
    utt = O.Utt(1, 1)
    # player 0: base (hp 10/10), worker carrying 1 resource; player 1: base with 4 of 10 hit points; 5 resources each
    g = O.Game(utt, _tiny_map([("Base", 0, 1, 1, 0, 10), ("Worker", 0, 2, 2, 1, 1), ("Base", 1, 6, 6, 0, 4)]))
    # Sqrt3: s0 = 5*20 + (1*10 + 40*1*sqrt(1/1)) + 40*10*sqrt(10/10) = 550 ; s1 = 5*20 + 40*10*sqrt(4/10 = 0) = 100
    want = np.float32(2 * np.float32(550) / np.float32(650)) - np.float32(1)
    assert np.float32(g.evaluate(0, 0, 1)) == np.float32(want)
    assert np.float32(g.evaluate(0, 1, 0)) == np.float32(np.float32(2 * np.float32(100) / np.float32(650)) - np.float32(1))
    # Simple: s0 = 100 + 10 + 40*1*1/1 + 40*10*10/10 = 550 ; s1 = 100 + 40*10*4/10 = 260
    assert g.evaluate(1, 0, 1) == 290.0 and g.evaluate(1, 1, 0) == -290.0
    # a player without units scores 0 in Sqrt3 (:41-43), so the other side's evaluation is 2*s/(s+0) - 1 = 1
    g = O.Game(utt, _tiny_map([("Base", 0, 1, 1, 0, 10)]))
    assert g.evaluate(0, 0, 1) == 1.0 and g.evaluate(0, 1, 0) == -1.0


def test_vector_action_known_answers(maps):
    """PlayerAction.fromVectorAction (PlayerAction.java:384-417) + UnitAction.fromVectorAction (UnitAction.java:675-709) on the
    initial state of maps/8x8/basesWorkers8x8.xml: rows are [cell, type, moveDir, harvestDir, returnDir, produceDir, produceType,
    attackIdx]; rows that name an empty cell, the opponent's unit or a cell already used by an earlier row are dropped."""
    utt = O.Utt(1, 1)
    g = O.Game(utt, maps["8x8/basesWorkers8x8"])
    rows = np.array([
        [1 + 1 * 8, O.MOVE, 2, 0, 0, 0, 0, 0],      # worker (1,1): move DOWN (to (1,2))
        [2 + 1 * 8, O.PRODUCE, 0, 0, 0, 2, 3, 0],   # base (2,1): produce a Worker (type 3) DOWN (to (2,2))
        [6 + 6 * 8, O.MOVE, 0, 0, 0, 0, 0, 0],      # the opponent's worker: ignored
        [3 + 3 * 8, O.MOVE, 1, 0, 0, 0, 0, 0],      # empty cell: ignored
    ], dtype=np.int32)
    pa = g.from_vector_action(0, rows, fill_none=1)
    # unit-list indices: 4 = worker p0, 2 = base p0; action = (type, parameter, x, y, unit type)
    assert pa == [(4, (O.MOVE, 2, 0, 0, -1)), (2, (O.PRODUCE, 2, 0, 0, 3))]
    # two rows whose actions target the same cell: the second is inconsistent with the resource usage built so far and is dropped
    rows = np.array([[2 + 1 * 8, O.PRODUCE, 0, 0, 0, 3, 3, 0],   # base (2,1) produces LEFT?  (1,1) is the worker's cell -- still a position
                     [1 + 1 * 8, O.MOVE, 1, 0, 0, 0, 0, 0]], dtype=np.int32)
    pa = g.from_vector_action(0, rows, fill_none=1)
    assert pa[0] == (2, (O.PRODUCE, 3, 0, 0, 3))
    # an attack index is relative to the unit on a (2a+1)^2 window, a = the largest attack range of the table (3 -> 7x7):
    # index 3 + 2*7 = 17 is (dx, dy) = (0, -1)
    rows = np.array([[1 + 1 * 8, O.ATTACK, 0, 0, 0, 0, 0, 17]], dtype=np.int32)
    pa = g.from_vector_action(0, rows, fill_none=1)
    # the base was not addressed: JNIAI.getAction fills it with NONE of duration 1 (ai/jni/JNIAI.java:53)
    assert pa == [(4, (O.ATTACK, -1, 1, 0, -1)), (2, (O.NONE, 1, 0, 0, -1))]


class _JavaRandom:
    """java.util.Random as the Java SE specification defines it (48-bit LCG), written independently of the oracle."""
    def __init__(self, seed):
        self.s = (seed ^ 0x5DEECE66D) & ((1 << 48) - 1)

    def next(self, bits):
        self.s = (self.s * 0x5DEECE66D + 0xB) & ((1 << 48) - 1)
        return self.s >> (48 - bits)

    def next_double(self):
        return ((self.next(26) << 27) + self.next(27)) * (1.0 / (1 << 53))


def test_random_biased_known_answer(maps):
    """RandomBiasedAI.getAction (RandomBiasedAI.java:51-107) + Sampler.weighted (Sampler.java:116-137) for player 0 on the initial
    state of maps/8x8/basesWorkers8x8.xml with util.Sampler.generator = new Random(42), action lists written out by hand from
    Unit.getUnitActions (Unit.java:382-522)."""
    r = _JavaRandom(42)
    d1, d2 = r.next_double(), r.next_double()
    assert d1 == 0.7275636800328681                      # the known answer of the Java SE specification
    # base (2,1), 5 resources: Worker (cost 1) up / right / down (the worker is on its left), then NONE: four actions of weight 1
    base_list = [(O.PRODUCE, 0, 3), (O.PRODUCE, 1, 3), (O.PRODUCE, 2, 3), (O.NONE, 10, -1)]
    # worker (1,1): Base (10) is too expensive; Barracks (5) up / down / left (the base is on its right); move up / down / left; NONE
    worker_list = [(O.PRODUCE, 0, 2), (O.PRODUCE, 2, 2), (O.PRODUCE, 3, 2), (O.MOVE, 0, -1), (O.MOVE, 2, -1), (O.MOVE, 3, -1), (O.NONE, 10, -1)]

    def weighted(draw, n):  # all weights are 1 here: first i with i + 1 >= draw * n
        tmp, acc = draw * n, 0.0
        for i in range(n):
            acc += 1.0
            if acc >= tmp:
                return i
        return n - 1

    b = base_list[weighted(d1, len(base_list))]
    w = worker_list[weighted(d2, len(worker_list))]
    g = O.Game(O.Utt(1, 1), maps["8x8/basesWorkers8x8"])
    g.seed(42)
    pa = g.random_biased(0)
    # units are visited in unit-list order: the base (index 2) before the worker (index 4); the two choices use different cells
    assert pa == [(2, (b[0], b[1], 0, 0, b[2])), (4, (w[0], w[1], 0, 0, w[2]))]
    assert b == (O.PRODUCE, 2, 3) and w[0] in (O.MOVE, O.PRODUCE)   # 0.7275.. * 4 = 2.91 -> third action: a Worker below the base


def test_issue_conflict_known_answers():
    """GameState.issue (GameState.java:249-328) and issueSafe (:338-408) on positions small enough to follow by hand: two units that
    choose the same cell in the same cycle, under each move-conflict strategy, and an illegal action."""
    # a Light of player 0 at (1,1) (move time 8) and a Worker of player 1 at (3,1) (move time 10) both step into (2,1)
    units = [("Light", 0, 1, 1, 0, 4), ("Worker", 1, 3, 1, 0, 1), ("Base", 0, 0, 7, 0, 10), ("Base", 1, 7, 7, 0, 10)]
    move_right, move_left = (O.MOVE, 1, 0, 0, -1), (O.MOVE, 3, 0, 0, -1)
    # CANCEL_BOTH (1): both become NONE of duration min(8, 10) (:276-278, :289-296)
    g = O.Game(O.Utt(1, 1), _tiny_map(units))
    g.issue([(0, move_right)]); g.issue([(1, move_left)])
    a = g.assignments()
    assert a[0, :3].tolist() == [1, O.NONE, 8] and a[1, :3].tolist() == [1, O.NONE, 8]
    # CANCEL_ALTERNATING (3): the counter starts even, so the NEW action is the one cancelled (:283-286); the old MOVE stays
    g = O.Game(O.Utt(1, 3), _tiny_map(units))
    g.issue([(0, move_right)]); g.issue([(1, move_left)])
    a = g.assignments()
    assert a[0, :3].tolist() == [1, O.MOVE, 1] and a[1, :3].tolist() == [1, O.NONE, 8]
    # ... and the next conflict cancels the OLD one
    units2 = units + [("Light", 0, 1, 4, 0, 4), ("Worker", 1, 3, 4, 0, 1)]
    g = O.Game(O.Utt(1, 3), _tiny_map(units2))
    g.issue([(0, move_right), (4, move_right)]); g.issue([(1, move_left), (5, move_left)])
    a = g.assignments()
    assert a[0, :3].tolist() == [1, O.MOVE, 1] and a[1, :3].tolist() == [1, O.NONE, 8]
    assert a[4, :3].tolist() == [1, O.NONE, 8] and a[5, :3].tolist() == [1, O.MOVE, 3]
    # after 8 cycles the surviving MOVE of the Light has executed
    g = O.Game(O.Utt(1, 3), _tiny_map(units))
    g.issue([(0, move_right)]); g.issue([(1, move_left)])
    for _ in range(8):
        g.cycle()
    u = g.units()
    assert (int(u[0, 2]), int(u[0, 3])) == (2, 1) and (int(u[1, 2]), int(u[1, 3])) == (3, 1)
    # issueSafe: an action that is not among the unit's legal actions is replaced by NONE with the action's own duration (:352-366):
    # the base cannot move
    g = O.Game(O.Utt(1, 1), _tiny_map(units))
    g.issue([(2, (O.MOVE, 0, 0, 0, -1))], safe=True)
    a = g.assignments()
    assert a[2, 0] == 1 and a[2, 1] == O.NONE


def test_execute_known_answers():
    """UnitAction.execute (UnitAction.java:338-465) and ETA (:307-329) followed by hand with the VERSION_ORIGINAL numbers of
    UnitTypeTable.java (worker: harvest 20, return = move time 10, attack 5, produced in 50; harvest amount 1; Light: 2 damage)."""
    utt = O.Utt(1, 1)
    # resource (0,0) with 20, worker p0 (1,0), base p0 (2,0); an enemy base far away keeps the game alive
    g = O.Game(utt, _tiny_map([("Resource", -1, 0, 0, 20, 1), ("Worker", 0, 1, 0, 0, 1), ("Base", 0, 2, 0, 0, 10), ("Base", 1, 7, 7, 0, 10)]))
    g.issue([(1, (O.HARVEST, 3, 0, 0, -1))])                 # harvest LEFT
    for _ in range(19):
        g.cycle()
    assert int(g.units()[1, 4]) == 0                         # not yet: ETA 20
    g.cycle()
    u = g.units()
    assert int(u[0, 4]) == 19 and int(u[1, 4]) == 1          # the resource lost one unit, the worker carries it
    g.issue([(1, (O.RETURN, 1, 0, 0, -1))])                  # return RIGHT to the base: ETA = move time 10 (:316-318)
    for _ in range(10):
        g.cycle()
    assert [g.resources(0), g.resources(1)] == [6, 5] and int(g.units()[1, 4]) == 0
    # the base trains a Worker DOWN: the cost leaves the player when the action EXECUTES, 50 cycles later (:404-418), and the new unit
    # joins the END of the unit list
    g.issue([(2, (O.PRODUCE, 2, 0, 0, 3))])
    for _ in range(49):
        g.cycle()
    assert [g.resources(0), g.resources(1)] == [6, 5] and g.n_units == 4
    g.cycle()
    u = g.units()
    assert [g.resources(0), g.resources(1)] == [5, 5] and g.n_units == 5 and u[4, :4].tolist() == [3, 0, 2, 1]
    # a Light (2 damage) attacks the adjacent enemy Worker (1 hp): after attackTime 5 the worker is gone, and with it the assignment it had
    g = O.Game(utt, _tiny_map([("Light", 0, 1, 1, 0, 4), ("Worker", 1, 2, 1, 0, 1), ("Base", 0, 0, 7, 0, 10), ("Base", 1, 7, 7, 0, 10)]))
    g.issue([(0, (O.ATTACK, -1, 2, 1, -1))]); g.issue([(1, (O.MOVE, 1, 0, 0, -1))])
    for _ in range(5):
        g.cycle()
    u = g.units()
    assert g.n_units == 3 and [int(t) for t in u[:, 0]] == [4, 1, 1] and g.assignments()[:, 0].tolist() == [0, 0, 0]


def test_worker_rush_first_decision_known_answer(maps):
    """WorkerRush.getAction (WorkerRush.java:63-204) on the initial state of maps/8x8/basesWorkers8x8.xml, by hand:
    the base trains a Worker into the free neighbour closest to a resource (Train.java:48-128: up (2,0), distance 2, beats right and
    down at 4); the worker becomes the harvester and A* to a cell adjacent to the resource at (0,0) expands up (1,0), down, left
    (0,1) and pops the NEWEST node of the smallest f first (AStarPathFinding.java:104-138) -- (0,1), already adjacent: first move LEFT.
    BFS pops the oldest, (1,0): UP.  Player 1 is the mirror image: train DOWN (5,7), move DOWN (A*) / RIGHT (BFS)."""
    g = O.Game(O.Utt(1, 1), maps["8x8/basesWorkers8x8"])
    assert O.ScriptedAI(O.AI_WORKER_RUSH, O.PF_ASTAR).get_action(g, 0) == [(2, (O.PRODUCE, 0, 0, 0, 3)), (4, (O.MOVE, 3, 0, 0, -1))]
    assert O.ScriptedAI(O.AI_WORKER_RUSH, O.PF_BFS).get_action(g, 0) == [(2, (O.PRODUCE, 0, 0, 0, 3)), (4, (O.MOVE, 0, 0, 0, -1))]
    assert O.ScriptedAI(O.AI_WORKER_RUSH, O.PF_ASTAR).get_action(g, 1) == [(3, (O.PRODUCE, 2, 0, 0, 3)), (5, (O.MOVE, 2, 0, 0, -1))]
    assert O.ScriptedAI(O.AI_WORKER_RUSH, O.PF_BFS).get_action(g, 1) == [(3, (O.PRODUCE, 2, 0, 0, 3)), (5, (O.MOVE, 1, 0, 0, -1))]


def _po_expectations(o):
    """player 0's 8-plane observation of the initial state of maps/8x8/basesWorkers8x8.xml (PartiallyObservableGameState.java:82-179):
    its base (2,1) sees 5 cells far, its worker (1,1) 3; nothing of player 1 is in sight, nor the resource at (7,7)."""
    assert o.shape == (8, 8, 8)
    assert o[3, 1, 2] == 2 and o[3, 1, 1] == 4 and o[3, 0, 0] == 1       # own base, own worker, the near resource (type + 1)
    assert o[3, 6, 5] == 0 and o[3, 6, 6] == 0 and o[3, 7, 7] == 0       # the enemy base and worker and the far resource are hidden
    assert o[0, 6, 5] == 0 and o[1, 7, 7] == 0 and o[2, 6, 6] == 0
    vis = o[6]
    assert vis[1, 7] == 1 and vis[2, 7] == 0      # (7,1): dx = 5 from the base, 25 <= 25; (7,2): 26
    assert vis[6, 2] == 1 and vis[7, 2] == 0      # (2,6): dy = 5; (2,7): 6
    assert vis[4, 6] == 1 and vis[5, 6] == 0      # (6,4): 16 + 9 = 25; (6,5): 16 + 16 = 32
    assert vis[0, 0] == 1 and vis[7, 7] == 0
    assert not o[7].any()                         # no enemy unit is visible, so nothing is known about what the opponent sees


def test_partially_observable_observation_known_answers(maps):
    g = O.Game(O.Utt(1, 1), maps["8x8/basesWorkers8x8"])
    _po_expectations(g.po_view(0).observe(0, po=True))
