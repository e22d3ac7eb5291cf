"""Host side of the search AIs over the batched engine against the oracle's restatements (parity unpinned: the reference holds no
golden data for them and its generators are unseeded):

  GameState.getPlayerActions            src/rts/GameState.java:493-524
  PlayerActionGenerator                 src/rts/PlayerActionGenerator.java:56-252
  NaiveMCTS / NaiveMCTSNode             src/ai/mcts/naivemcts/*.java (seeded generators: the trees must come out node for node)
"""
import numpy as np
import pytest

import microrts_b200 as M
import parity as P
from microrts_b200 import search as S
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def rows_of(pairs, units, w):
    """oracle (unit index, action) pairs -> RAW rows like the library's"""
    return [[int(units[ui][2] + units[ui][3] * w), ty, par, (x if ty == O.ATTACK else 0), (y if ty == O.ATTACK else 0), (ut if ty == O.PRODUCE else -1), 0, 0]
            for (ui, (ty, par, x, y, ut)) in pairs]


def advanced_games(backend, maps, key, n, version=1, warm=None):
    utt, outt = M.UnitTypeTable(version, 1), O.Utt(version, 1)
    b = M.BatchedGameState(utt, M.maps.standard_map(key, utt), n)
    seeds = np.arange(n, dtype=np.int64) * 13 + 5
    b.reset(seeds)
    b.set_policy(0, M.POLICY_RANDOM_BIASED)
    b.set_policy(1, M.POLICY_RANDOM_BIASED)
    games = []
    for g in range(n):
        og = O.Game(outt, maps[key])
        og.seed(int(seeds[g]))
        games.append(og)
    warm = warm or [37 * (g + 1) for g in range(n)]
    tmp = M.BatchedGameState(utt, M.maps.standard_map(key, utt), n)
    tmp.set_policy(0, M.POLICY_RANDOM_BIASED)
    tmp.set_policy(1, M.POLICY_RANDOM_BIASED)
    for g in range(n):  # game g advanced by warm[g] cycles
        tmp.copy_games(b)
        tmp.step(warm[g], 3000)
        b.copy_games(tmp, mask=(np.arange(n) == g).astype(np.uint8))
        games[g].run(O.AI_RANDOM_BIASED, None, O.AI_RANDOM_BIASED, None, warm[g], 3000)
    tmp.close()
    b.cycle_to_decision()  # states in which somebody can act (after a full RandomBiasedAI cycle every unit is busy)
    for og in games:
        while og.winner == -1 and not og.gameover and og.is_complete():
            og.cycle()
    ex = b.export()
    for g in range(n):
        P.assert_same_state(ex, g, games[g], "warm")
    return utt, b, games


@pytest.mark.parametrize("key", ["8x8/basesWorkers8x8", "16x16/basesWorkers16x16"])
def test_player_actions_and_generator(backend, maps, key):
    n = 3 if backend == "emu" else 6
    utt, b, games = advanced_games(backend, maps, key, n)
    w = maps[key]["w"]
    checked = 0
    for g, og in enumerate(games):
        units = og.units()
        for player in (0, 1):
            asg = og.assignments()
            if not any(units[i][1] == player and not asg[i][0] for i in range(len(units))):
                with pytest.raises(M.MicroRTSError):
                    S.PlayerActionGenerator(b, g, player)
                continue
            gen, ogen = S.PlayerActionGenerator(b, g, player), O.Pag(og, player)
            assert gen.getSize() == ogen.size
            if ogen.size <= 200000:
                ref, total = O.player_actions(og, player)
                got, gtotal = S.player_actions(b, g, player, max_player_actions=max(1, total))
                assert gtotal == total == len(ref)
                for a, r in zip(got, ref):
                    assert a.tolist() == rows_of(r, units, w)
                k = 0
                while True:  # the odometer: every consistent combination, last choice first
                    r = ogen.next()
                    a = gen.getNextAction()
                    if r is None:
                        assert a is None
                        break
                    assert a.tolist() == rows_of(r, units, w), (key, g, player, k)
                    k += 1
                assert k == total and gen.getGenerated() == ogen.generated == total
                checked += 1
            gen, ogen = S.PlayerActionGenerator(b, g, player), O.Pag(og, player)
            rng, orng = S.JavaRandomState(99 + g), O.JavaRandom(99 + g)
            for _ in range(5):
                assert gen.getRandom(rng).tolist() == rows_of(ogen.random(orng), units, w)
            gen.randomizeOrder(rng); ogen.randomize_order(orng)
            for _ in range(10):
                r, a = ogen.next(), gen.getNextAction()
                assert (a is None) == (r is None)
                if r is None:
                    break
                assert a.tolist() == rows_of(r, units, w)
    assert checked > 0
    b.close()


@pytest.mark.parametrize("key,player,strategy,eval_fn", [("8x8/basesWorkers8x8", 0, 0, 0), ("8x8/basesWorkers8x8", 1, 1, 1), ("16x16/basesWorkers16x16", 0, 0, 0),
                                                          ("melee14x12Mixed18", 1, 0, 0)])
def test_naive_mcts_trees_equal_the_oracle(backend, maps, key, player, strategy, eval_fn):
    n = 3 if backend == "emu" else 12
    iters = 40 if backend == "emu" else 300
    utt, b, games = advanced_games(backend, maps, key, n, warm=[0] + [29 * g + 11 for g in range(1, n)])
    seeds = np.arange(n, dtype=np.int64) * 7 + 1
    before = b.export()
    search = S.NaiveMCTS(b, player, seeds=seeds, lookahead=100, max_depth=10, epsilon_l=0.3, epsilon_g=0.0, epsilon_0=0.4, global_strategy=strategy,
                         eval_fn=eval_fn, max_nodes_per_tree=iters + 2)
    refs = [O.Mcts(og, player, int(seeds[g]), 100, 10, 0.3, 0.0, 0.4, strategy, True, eval_fn) for g, og in enumerate(games)]
    done = 0
    for chunk in (1, 7, iters - 8):
        search.iterate(chunk)
        done += chunk
        for g, ref in enumerate(refs):
            ref.iterate(chunk)
            rv, ra, cv, ca = search.root(g)
            orv, ora, ocv, oca = ref.root()
            assert rv == orv == done and search.num_nodes(g) == ref.n_nodes, (key, g, done, search.num_nodes(g), ref.n_nodes)
            assert (cv == ocv).all() and (ca == oca).all() and ra == ora, "tree %d after %d iterations\nvisits %s\noracle %s" % (g, done, cv, ocv)
    rows, counts = search.best_actions()
    w = maps[key]["w"]
    for g, ref in enumerate(refs):
        best = ref.best_action()
        exp = [] if best is None else rows_of(best, games[g].units(), w)
        assert rows[g, :counts[g]].tolist() == exp, (key, g)
    after = b.export()
    for k in ("header", "units", "actions", "rng"):
        assert (before[k] == after[k]).all(), "the search modified the root batch"
    search.close()
    b.close()


@pytest.mark.parametrize("key,player", [("8x8/basesWorkers8x8", 0), ("16x16/basesWorkers16x16", 1)])
def test_uct_trees_equal_the_oracle(backend, maps, key, player):
    """ai.mcts.uct.UCT: every node's shuffled move generator is exhausted before UCB1 chooses among its children."""
    n = 3 if backend == "emu" else 10
    iters = 40 if backend == "emu" else 250
    utt, b, games = advanced_games(backend, maps, key, n, warm=[0] + [31 * g + 7 for g in range(1, n)])
    seeds = np.arange(n, dtype=np.int64) * 11 + 3
    search = S.UCT(b, player, seeds=seeds, lookahead=100, max_depth=10, max_nodes_per_tree=iters + 2)
    refs = [O.Uct(og, player, int(seeds[g]), 100, 10, 0) for g, og in enumerate(games)]
    done = 0
    for chunk in (1, 9, iters - 10):
        search.iterate(chunk)
        done += chunk
        for g, ref in enumerate(refs):
            ref.iterate(chunk)
            rv, ra, cv, ca = search.root(g)
            orv, ora, ocv, oca = ref.root()
            assert rv == orv == done and search.num_nodes(g) == ref.n_nodes, (key, g, done, search.num_nodes(g), ref.n_nodes)
            assert (cv == ocv).all() and (ca == oca).all() and ra == ora, "tree %d after %d iterations\nvisits %s\noracle %s" % (g, done, cv, ocv)
    rows, counts = search.best_actions()
    w = maps[key]["w"]
    for g, ref in enumerate(refs):
        best = ref.best_action()
        exp = [] if best is None else rows_of(best, games[g].units(), w)
        assert rows[g, :counts[g]].tolist() == exp, (key, g)
    search.close()
    b.close()


def test_many_searches_in_lockstep(backend, maps):
    """160 searches at once (the host side of the searches runs on several threads above 128): a sample of the trees equals the oracle."""
    key, n, iters = "8x8/basesWorkers8x8", 160, (6 if backend == "emu" else 60)
    utt, base, games4 = advanced_games(backend, maps, key, 4, warm=[0, 40, 90, 150])
    b = M.BatchedGameState(utt, M.maps.standard_map(key, utt), n)
    b.copy_games(base, src_index=np.arange(n, dtype=np.int64) % 4)
    seeds = np.arange(n, dtype=np.int64) * 3 + 1000
    search = S.NaiveMCTS(b, 0, seeds=seeds, max_nodes_per_tree=iters + 2)
    search.iterate(iters)
    for g in (0, 1, 2, 3, 77, 130, 159):
        ref = O.Mcts(games4[g % 4], 0, int(seeds[g]))
        ref.iterate(iters)
        rv, ra, cv, ca = search.root(g)
        orv, ora, ocv, oca = ref.root()
        assert rv == orv == iters and search.num_nodes(g) == ref.n_nodes and (cv == ocv).all() and (ca == oca).all() and ra == ora, g
    search.close()
    b.close()
    base.close()
