"""JNIGridnetVecClient (microrts_b200.vec_client) against a restatement of the reference's client classes over the CPU oracle.

RefSelfPlay restates JNIGridnetClientSelfPlay.gameStep (src/tests/JNIGridnetClientSelfPlay.java:157-190), RefBotEnv restates
JNIGridnetClient.gameStep (src/tests/JNIGridnetClient.java:163-203), ref_rewards restates src/ai/reward/*.java from the
TraceEntry contents (unit list before the cycle + the PlayerActions as issueSafe left them) and RefVec restates the
auto-reset of JNIGridnetVecClient.gameStep (src/tests/JNIGridnetVecClient.java:213-297).  Parity unpinned: the reference
holds no golden data for this facade; the oracle follows the cited files.
"""
import math
import os

import numpy as np
import pytest

import microrts_b200 as M
import parity as P
from microrts_b200 import rewards as R
from microrts_b200.vec_client import JNIGridnetVecClient, ai
from oracle import oracle as O

pytestmark = pytest.mark.gpu

RFS = [R.WinLossRewardFunction, R.ResourceGatherRewardFunction, R.ProduceWorkerRewardFunction, R.ProduceBuildingRewardFunction,
       R.AttackRewardFunction, R.ProduceCombatUnitRewardFunction, R.CloserToEnemyBaseRewardFunction]
NAMES = O.TYPE_NAMES


def ref_rewards(maxplayer, pre_units, pairs_by_player, og):
    """[(reward, done)] in RFS order, computed the way the Java classes do."""
    minplayer = 1 - maxplayer
    post = og.units()
    mine = pairs_by_player[maxplayer]
    out = []
    # WinLoss
    out.append(((1.0 if og.winner == maxplayer else -1.0), True) if og.gameover else (0.0, False))
    # ResourceGather
    r = sum(1.0 for (_u, a) in mine if a[0] == O.HARVEST) + sum(1.0 for (_u, a) in mine if a[0] == O.RETURN)
    out.append((r, not any(NAMES[u[0]] == "Resource" and u[4] > 0 for u in post)))
    # ProduceWorker / ProduceBuilding
    out.append((sum(1.0 for (_u, a) in mine if a[0] == O.PRODUCE and a[4] >= 0 and NAMES[a[4]] == "Worker"), False))
    out.append((sum(1.0 for (_u, a) in mine if a[0] == O.PRODUCE and a[4] >= 0 and NAMES[a[4]] in ("Barracks", "Base")), False))
    # Attack: the unit at the attacked cell in the TraceEntry's PhysicalGameState
    r = 0.0
    for (_u, a) in mine:
        if a[0] == O.ATTACK:
            other = [u for u in pre_units if u[2] == a[2] and u[3] == a[3]]
            if other:
                if other[0][1] == minplayer:
                    r += 1.0
                elif other[0][1] == maxplayer:
                    r -= 1.0
    out.append((r, False))
    out.append((sum(1.0 for (_u, a) in mine if a[0] == O.PRODUCE and a[4] >= 0 and NAMES[a[4]] in ("Light", "Heavy", "Ranged")), False))
    # CloserToEnemyBase
    base = [u for u in pre_units if u[1] == minplayer and NAMES[u[0]] == "Base"]
    if not base:
        out.append((0.0, False))
    else:
        bx, by = base[0][2], base[0][3]

        def closest(units):
            d = 2000000000.0
            for u in units:
                if u[1] == maxplayer and NAMES[u[0]] in ("Light", "Heavy", "Ranged", "Worker"):
                    d = min(d, math.sqrt(math.pow(bx - u[2], 2.0) + math.pow(by - u[3], 2.0)))
            return d
        out.append((closest(pre_units) - closest(post), False))
    return out


class RefEnvBase:
    def __init__(self, outt, mapd, seed):
        self.outt, self.mapd = outt, mapd
        self.og = O.Game(outt, mapd)
        self.og.seed(seed)

    def restart(self):
        old = self.og
        self.og = O.Game(self.outt, self.mapd)
        for k in range(3):  # the reference's Random objects are static: they keep running across resets
            self.og.set_rng_state(k, old.rng_state(k))
        self.on_restart()

    def on_restart(self):
        pass


class RefSelfPlay(RefEnvBase):
    def step(self, a0, a1):
        og = self.og
        pre = [tuple(u) for u in og.units()]
        pas = []
        for pl, act in ((0, a0), (1, a1)):
            pa = og.from_vector_action(pl, act, fill_none=1)
            pas.append(og.issue_out(pa, safe=True))
        og.cycle()
        return [ref_rewards(pl, pre, pas, og) for pl in (0, 1)]


def view_pairs_to_game(view, og, pairs):
    """unit indices of a PartiallyObservableGameState's list -> indices of the real state's list (same units, found by position)"""
    vu, gu = view.units(), og.units()
    at = {(int(u[2]), int(u[3])): i for i, u in enumerate(gu)}
    return [(at[(int(vu[i][2]), int(vu[i][3]))], a) for i, a in pairs]


class RefBotEnv(RefEnvBase):
    def __init__(self, outt, mapd, seed, kind, pf, side, partial=False):
        super().__init__(outt, mapd, seed)
        self.kind, self.pf, self.side, self.partial = kind, pf, side, partial
        self.on_restart()

    def on_restart(self):
        self.bot = O.ScriptedAI(self.kind, self.pf) if self.kind in O.SCRIPTED_AIS else None

    def step(self, act):
        og, pl = self.og, self.side
        pre = [tuple(u) for u in og.units()]
        if self.partial:
            # JNIGridnetClient.gameStep :163-172: both players decide on their PartiallyObservableGameState, issueSafe on the real state
            v1, v2 = og.po_view(pl), og.po_view(1 - pl)
            pa1 = view_pairs_to_game(v1, og, v1.from_vector_action(pl, act, fill_none=1))
            if self.kind == O.AI_RANDOM_BIASED:
                pa2 = view_pairs_to_game(v2, og, v2.random_biased(1 - pl))
                for k in range(3):  # RandomBiasedAI's generator is not part of the view
                    og.set_rng_state(k, v2.rng_state(k))
            elif self.bot is not None:
                pa2 = view_pairs_to_game(v2, og, self.bot.get_action(v2, 1 - pl))
            else:
                pa2 = []
        else:
            pa1 = og.from_vector_action(pl, act, fill_none=1)
            if self.kind == O.AI_RANDOM_BIASED:
                pa2 = og.random_biased(1 - pl)
            elif self.bot is not None:
                pa2 = self.bot.get_action(og, 1 - pl)
            else:
                pa2 = []
        pas = [None, None]
        pas[pl] = og.issue_out(pa1, safe=True)
        pas[1 - pl] = og.issue_out(pa2, safe=True)
        og.cycle()
        return ref_rewards(pl, pre, pas, og)


def random_actions(rng, og, player, w, h, k):
    """Vector-action rows for `player`: mostly rows on own idle units with plausible parameters, some noise."""
    rows = np.zeros((k, 8), dtype=np.int32)
    units = og.units()
    asg = og.assignments()
    own = [i for i in range(len(units)) if units[i][1] == player and asg[i][0] == 0]
    rng.shuffle(own)
    for r in range(k):
        if r < len(own) and rng.random() < 0.9:
            u = units[own[r]]
            cell = int(u[2] + u[3] * w)
        else:
            cell = int(rng.integers(0, w * h))
        rows[r] = [cell, int(rng.choice([0, 1, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5])), rng.integers(0, 4), rng.integers(0, 4), rng.integers(0, 4),
                   rng.integers(0, 4), int(rng.choice([1, 2, 3, 3, 3, 4, 5, 6])), rng.integers(0, 49)]
    return rows


@pytest.mark.parametrize("key,n_sp,bots,max_steps,compact,partial", [
    ("8x8/basesWorkers8x8", 4, ["RandomBiasedAI", "PassiveAI", "WorkerRush"], 60, False, False),
    ("16x16/basesWorkers16x16", 2, ["LightRush", "RandomBiasedAI"], 150, False, False),
    ("8x8/basesWorkers8x8", 6, [], 45, True, False),                 # one self-play group: the client's arrays are the group's pinned arrays
    ("16x16/basesWorkers16x16", 2, ["WorkerRush", "PassiveAI"], 80, True, False),
    # partial_obs = true: 8-plane observations of the agent's view, the opponent decides on ITS view.  (The reference hands out the
    # view object built BEFORE the cycle -- stale unit list, live units; here the observation is of the state after the cycle:
    # DESIGN.md section 8, deviation 2.)
    ("16x16/basesWorkers16x16", 2, ["LightRush", "RandomBiasedAI", "WorkerRush"], 200, False, True),
])
def test_vec_client_matches_reference_flow(backend, maps, tmp_path, key, n_sp, bots, max_steps, compact, partial):
    m = maps[key]
    path = tmp_path / "map.xml"
    path.write_text(P.map_to_xml(m))
    utt, outt = M.UnitTypeTable(1, 1), O.Utt(1, 1)
    rfs = [c() for c in RFS]
    n_envs = len(bots)
    s1 = n_sp + n_envs
    specs = [getattr(ai, b)(utt) for b in bots]
    kinds = {"RandomBiasedAI": O.AI_RANDOM_BIASED, "PassiveAI": O.AI_PASSIVE, "WorkerRush": O.AI_WORKER_RUSH, "LightRush": O.AI_LIGHT_RUSH}
    seed = 123
    vc = JNIGridnetVecClient(n_sp, n_envs, max_steps, rfs, "", [str(path)] * s1, specs, utt, partial_obs=partial, seed=seed, compact=compact)
    players = [0] * s1
    resp = vc.reset(players)
    ref_sp = [RefSelfPlay(outt, m, seed + 2 * i) for i in range(n_sp // 2)]
    ref_bot = [RefBotEnv(outt, m, seed + n_sp + i, kinds[b], 0, 0, partial) for i, b in enumerate(bots)]
    steps = np.zeros(s1, dtype=int)

    def ref_obs(og, pl):
        return og.po_view(pl).observe(pl, po=True) if partial else og.observe(pl)

    def check_obs_masks(ctx):
        mk = vc.getMasks(0)
        if compact:  # uint8 observations; the bit-packed masks that came with them equal the dense ones
            assert resp.observation.dtype == np.uint8
            packed = vc.getMasksPacked()
            assert (np.unpackbits(packed, axis=-1, bitorder="little")[..., :mk.shape[-1]] == mk).all(), (ctx, "packed masks")
        for i, e in enumerate(ref_sp):
            for pl in (0, 1):
                assert (resp.observation[2 * i + pl] == ref_obs(e.og, pl)).all(), (ctx, "obs selfplay", i, pl)
                assert (mk[2 * i + pl] == e.og.masks(pl)).all(), (ctx, "masks selfplay", i, pl)
        for i, e in enumerate(ref_bot):
            assert (resp.observation[n_sp + i] == ref_obs(e.og, 0)).all(), (ctx, "obs bot", i)
            assert (mk[n_sp + i] == e.og.masks(0)).all(), (ctx, "masks bot", i)

    check_obs_masks("reset")
    rng = np.random.default_rng(5)
    K = 6
    total = int(os.environ.get("MRTS_VEC_STEPS", "40" if backend == "emu" else "400"))
    n_resets = 0
    for t in range(total):
        action = np.zeros((s1, K, 8), dtype=np.int32)
        for i, e in enumerate(ref_sp):
            action[2 * i] = random_actions(rng, e.og, 0, m["w"], m["h"], K)
            action[2 * i + 1] = random_actions(rng, e.og, 1, m["w"], m["h"], K)
        for i, e in enumerate(ref_bot):
            action[n_sp + i] = random_actions(rng, e.og, 0, m["w"], m["h"], K)
        resp = vc.gameStep(action, players)
        exp_r = np.zeros((s1, len(rfs)))
        exp_d = np.zeros((s1, len(rfs)), dtype=bool)
        for i, e in enumerate(ref_sp):
            both = e.step(action[2 * i], action[2 * i + 1])
            for pl in (0, 1):
                exp_r[2 * i + pl] = [x[0] for x in both[pl]]
                exp_d[2 * i + pl] = [x[1] for x in both[pl]]
            steps[2 * i] += 1
            if exp_d[2 * i, 0] or steps[2 * i] >= max_steps:
                e.restart(); steps[2 * i] = 0; exp_d[2 * i, 0] = exp_d[2 * i + 1, 0] = True; n_resets += 1
        for i, e in enumerate(ref_bot):
            rw = e.step(action[n_sp + i])
            exp_r[n_sp + i] = [x[0] for x in rw]
            exp_d[n_sp + i] = [x[1] for x in rw]
            steps[n_sp + i] += 1
            if exp_d[n_sp + i, 0] or steps[n_sp + i] >= max_steps:
                e.restart(); steps[n_sp + i] = 0; exp_d[n_sp + i, 0] = True; n_resets += 1
        assert (resp.reward == exp_r).all(), "rewards at step %d\nours=\n%s\nref=\n%s" % (t, resp.reward, exp_r)
        assert (resp.done == exp_d).all(), "dones at step %d\nours=\n%s\nref=\n%s" % (t, resp.done, exp_d)
        if t % 7 == 0 or t == total - 1:
            check_obs_masks("step %d" % t)
        else:
            for i, e in enumerate(ref_sp):
                assert (resp.observation[2 * i] == ref_obs(e.og, 0)).all(), ("obs", t, i)
    assert n_resets > 0 or backend == "emu"
    vc.close()
