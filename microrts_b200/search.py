"""Host side of the reference's search AIs over the batched engine.

  PlayerActionGenerator   <- rts.PlayerActionGenerator (src/rts/PlayerActionGenerator.java:56-252): one game's PlayerActions in the
                             reference's enumeration order, random PlayerActions, randomizeOrder
  player_actions          <- GameState.getPlayerActions (src/rts/GameState.java:493-524)
  NaiveMCTS               <- ai.mcts.naivemcts.NaiveMCTS (src/ai/mcts/naivemcts/NaiveMCTS.java), one search per game of a batch, in lockstep

PlayerActions are lists of RAW action rows {cell, type, parameter, x, y, unit type, 0, 0}, the form BatchedGameState.issue() takes.
The generators the reference leaves unseeded are seeded by the caller (JavaRandomState).
"""
import ctypes as C

import numpy as np

from . import _ffi
from .api import _check


class JavaRandomState:
    """The 48-bit state of a java.util.Random, owned by the caller and advanced by the library calls that draw from it."""

    def __init__(self, seed):
        self.state = C.c_int64(_ffi.lib().mrts_java_random_seed(seed))


def player_actions(batch, game, player, max_player_actions=100000, max_k=None):
    """GameState.getPlayerActions(player) of one game: (list of PlayerActions as [k][8] int32 row arrays, total count)."""
    max_k = max_k or batch.cap
    rows = np.zeros((max_player_actions, max_k, 8), dtype=np.int32)
    counts = np.zeros(max_player_actions, dtype=np.int32)
    total = C.c_int64(0)
    _check(_ffi.lib().mrts_batch_player_actions(batch._h, game, player, rows.ctypes.data, counts.ctypes.data, max_player_actions, max_k, C.byref(total)))
    n = min(total.value, max_player_actions)
    return [rows[i, :counts[i]].copy() for i in range(n)], total.value


class PlayerActionGenerator:
    def __init__(self, batch, game, player, none_duration=10):
        h = C.c_void_p()
        _check(_ffi.lib().mrts_pag_create(batch._h, game, player, none_duration, C.byref(h)))
        self._h, self._max_k = h, batch.cap

    def __del__(self):
        try:
            _ffi.lib().mrts_pag_destroy(self._h)
        except Exception:
            pass

    def getSize(self):
        return _ffi.lib().mrts_pag_size(self._h)

    def getGenerated(self):
        return _ffi.lib().mrts_pag_generated(self._h)

    def getNextAction(self):
        """The next PlayerAction ([k][8] rows, last choice first as the reference adds them), or None when all have been generated."""
        rows = np.zeros((self._max_k, 8), dtype=np.int32)
        n = _ffi.lib().mrts_pag_next(self._h, rows.ctypes.data, self._max_k)
        if n == -100:
            return None
        _check(n)
        return rows[:n]

    def getRandom(self, rng):
        rows = np.zeros((self._max_k, 8), dtype=np.int32)
        n = _check(_ffi.lib().mrts_pag_random(self._h, C.byref(rng.state), rows.ctypes.data, self._max_k))
        return rows[:n]

    def randomizeOrder(self, rng):
        _check(_ffi.lib().mrts_pag_randomize_order(self._h, C.byref(rng.state)))


class NaiveMCTS:
    """One NaiveMCTS search per game of `roots` (the searching player's move in each), advanced in lockstep."""

    def __init__(self, roots, player, seeds=None, lookahead=100, max_depth=10, epsilon_l=0.3, epsilon_g=0.0, epsilon_0=0.4, global_strategy=0,
                 force_exploration=True, eval_fn=0, max_nodes_per_tree=1001, algorithm=0):
        prm = _ffi.MctsParams(lookahead, max_depth, epsilon_l, epsilon_g, epsilon_0, global_strategy, 1 if force_exploration else 0, eval_fn, algorithm)
        s = None if seeds is None else np.ascontiguousarray(seeds, dtype=np.int64)
        assert s is None or len(s) == roots.n
        h = C.c_void_p()
        _check(_ffi.lib().mrts_mcts_create(roots._h, player, C.byref(prm), max_nodes_per_tree, None if s is None else s.ctypes.data, C.byref(h)))
        self._h, self.n, self._max_k = h, roots.n, roots.cap

    def close(self):
        if self._h is not None:
            _ffi.lib().mrts_mcts_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def iterate(self, n_iterations=1):
        _check(_ffi.lib().mrts_mcts_iterate(self._h, n_iterations))

    def num_nodes(self, tree):
        return _check(_ffi.lib().mrts_mcts_num_nodes(self._h, tree))

    def root(self, tree, max_children=4096):
        """(visits, accumulated evaluation, children visits, children accumulated evaluation) of the root of one search."""
        rv, ra = C.c_int32(0), C.c_double(0)
        cv, ca = np.zeros(max_children, dtype=np.int32), np.zeros(max_children, dtype=np.float64)
        n = _check(_ffi.lib().mrts_mcts_root(self._h, tree, C.byref(rv), C.byref(ra), cv.ctypes.data, ca.ctypes.data, max_children))
        return rv.value, ra.value, cv[:n].copy(), ca[:n].copy()

    def best_actions(self):
        """getBestActionSoFar of every search: ([n][max_k][8] RAW rows, [n] counts) -- issue them with BatchedGameState.issue(player, rows, counts)."""
        rows = np.zeros((self.n, self._max_k, 8), dtype=np.int32)
        counts = np.zeros(self.n, dtype=np.int32)
        _check(_ffi.lib().mrts_mcts_best_actions(self._h, rows.ctypes.data, counts.ctypes.data, self._max_k))
        return rows, counts


class UCT(NaiveMCTS):
    """ai.mcts.uct.UCT (src/ai/mcts/uct/UCT.java, UCTNode.java), one search per game: the same lockstep machinery with UCTSelectLeaf."""

    def __init__(self, roots, player, seeds=None, lookahead=100, max_depth=10, eval_fn=0, max_nodes_per_tree=1001):
        super().__init__(roots, player, seeds=seeds, lookahead=lookahead, max_depth=max_depth, eval_fn=eval_fn, max_nodes_per_tree=max_nodes_per_tree, algorithm=1)
