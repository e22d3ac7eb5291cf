"""ctypes wrapper around oracle/libmrts_oracle.so -- TEST INFRASTRUCTURE ONLY.

The oracle is the CPU restatement of the reference's rules (oracle/mrts_oracle.c).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

TYPE_NAMES = ["Resource", "Base", "Barracks", "Worker", "Light", "Heavy", "Ranged"]
NONE, MOVE, HARVEST, RETURN, PRODUCE, ATTACK = range(6)
(AI_NONE, AI_PASSIVE, AI_RANDOM_BIASED, AI_WORKER_RUSH, AI_LIGHT_RUSH, AI_HEAVY_RUSH, AI_RANGED_RUSH,
 AI_WORKER_DEFENSE, AI_LIGHT_DEFENSE, AI_HEAVY_DEFENSE, AI_RANGED_DEFENSE,
 AI_PO_WORKER_RUSH, AI_PO_LIGHT_RUSH, AI_PO_HEAVY_RUSH, AI_PO_RANGED_RUSH, AI_WORKER_RUSH_PP, AI_CRUSH_V1, AI_CRUSH_V2, AI_EMR_DETERMINISTICO) = range(19)
SCRIPTED_AIS = tuple(range(AI_WORKER_RUSH, AI_EMR_DETERMINISTICO + 1))
PF_ASTAR, PF_BFS, PF_GREEDY, PF_FLOODFILL = 0, 1, 2, 3


def build(force=False):
    so = os.path.join(_HERE, "libmrts_oracle.so")
    src = os.path.join(_HERE, "mrts_oracle.c")
    hdr = os.path.join(_HERE, "mrts_oracle.h")
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["gcc", "-O2", "-g", "-std=c11", "-fPIC", "-shared", "-o", so, src, "-lm"])
    return so


class ActionV(C.Structure):
    _fields_ = [("type", C.c_int), ("param", C.c_int), ("x", C.c_int), ("y", C.c_int), ("utype", C.c_int)]

    def tup(self):
        return (self.type, self.param, self.x, self.y, self.utype)


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        vp, i, i64 = C.c_void_p, C.c_int, C.c_int64
        pi32 = C.POINTER(C.c_int32)
        sig = {
            "o_utt_create": (vp, [i, i]), "o_utt_empty": (vp, [i]),
            "o_utt_set_type": (None, [vp, i, C.POINTER(C.c_int16), i, i, C.POINTER(C.c_uint8)]),
            "o_utt_field": (i, [vp, i, i]), "o_utt_flags": (i, [vp, i]), "o_utt_produces": (i, [vp, i, C.POINTER(C.c_uint8)]),
            "o_utt_max_attack_range": (i, [vp]), "o_utt_free": (None, [vp]),
            "o_game_create": (vp, [vp, i, i, C.POINTER(C.c_uint8), i, i]),
            "o_game_add_unit": (None, [vp, i, i64, i, i, i, i, i]),
            "o_game_clone": (vp, [vp]), "o_game_free": (None, [vp]), "o_game_seed": (None, [vp, i64]),
            "o_game_time": (i, [vp]), "o_game_n_units": (i, [vp]), "o_game_resources": (i, [vp, i]),
            "o_game_rng_state": (i64, [vp, i]), "o_game_set_rng_state": (None, [vp, i, i64]), "o_game_winner": (i, [vp]), "o_game_gameover": (i, [vp]), "o_game_errors": (i, [vp]),
            "o_game_units": (i, [vp, pi32]), "o_game_assignments": (i, [vp, pi32]),
            "o_game_cycle": (i, [vp]), "o_game_is_complete": (i, [vp]), "o_game_next_change_time": (i, [vp]),
            "o_game_issue": (i, [vp, i, pi32, C.POINTER(ActionV), i]),
            "o_game_issue_out": (i, [vp, i, pi32, C.POINTER(ActionV), i]),
            "o_unit_actions": (i, [vp, i, i, C.POINTER(ActionV), i]),
            "o_game_free_cell": (i, [vp, i, i]),
            "o_ai_random_biased": (i, [vp, i, pi32, C.POINTER(ActionV)]),
            "o_ai_create": (vp, [i, i]), "o_ai_clone": (vp, [vp]), "o_ai_free": (None, [vp]),
            "o_ai_get_action": (i, [vp, vp, i, pi32, C.POINTER(ActionV)]),
            "o_from_vector_action": (i, [vp, i, i, pi32, i, pi32, C.POINTER(ActionV)]),
            "o_pathfind": (i, [vp, i, i, i, i, i, pi32]),
            "o_ff_create": (vp, []), "o_ff_free": (None, [vp]), "o_ff_find": (i, [vp, vp, i, i, i, i, pi32]),
            "o_observe": (None, [vp, i, pi32]), "o_observe_po": (None, [vp, i, pi32]), "o_masks": (None, [vp, i, pi32]),
            "o_po_view": (vp, [vp, i]),
            "o_evaluate": (C.c_float, [vp, i, i, i]),
            "o_run_game": (i, [vp, i, vp, i, vp, i, i, C.POINTER(C.c_int64)]),
            "o_simulate": (i, [vp, i]),
            "o_run_game_observing": (i, [vp, i, vp, i, vp, i, i, pi32]),
            "o_run_game_po": (i, [vp, i, vp, i, vp, i, i]),
            "o_pag_create": (vp, [vp, i, i]), "o_pag_free": (None, [vp]), "o_pag_size": (i64, [vp]), "o_pag_generated": (i64, [vp]), "o_pag_n_choices": (i, [vp]),
            "o_pag_next": (i, [vp, pi32, C.POINTER(ActionV)]), "o_pag_randomize_order": (None, [vp, C.POINTER(C.c_uint64)]),
            "o_pag_random": (i, [vp, C.POINTER(C.c_uint64), pi32, C.POINTER(ActionV)]), "o_player_actions": (i64, [vp, i, pi32, i64]),
            "o_mcts_create": (vp, [vp, i, i, i, C.c_float, C.c_float, C.c_float, i, i, i, i64]), "o_mcts_free": (None, [vp]), "o_mcts_iterate": (None, [vp, i]),
            "o_mcts_root": (i, [vp, pi32, C.POINTER(C.c_double), pi32, C.POINTER(C.c_double), i]), "o_mcts_n_nodes": (i, [vp]),
            "o_mcts_best_action": (i, [vp, pi32, C.POINTER(ActionV)]),
            "o_uct_create": (vp, [vp, i, i, i, i, i64]), "o_uct_free": (None, [vp]), "o_uct_iterate": (None, [vp, i]),
            "o_uct_root": (i, [vp, pi32, C.POINTER(C.c_float), pi32, C.POINTER(C.c_float), i]), "o_uct_n_nodes": (i, [vp]),
            "o_uct_best_action": (i, [vp, pi32, C.POINTER(ActionV)]),
            "o_jr_seed": (None, [C.POINTER(C.c_uint64), i64]), "o_jr_next": (C.c_int32, [C.POINTER(C.c_uint64), i]),
            "o_jr_next_int": (C.c_int32, [C.POINTER(C.c_uint64)]),
            "o_jr_next_int_bound": (C.c_int32, [C.POINTER(C.c_uint64), C.c_int32]),
            "o_jr_next_double": (C.c_double, [C.POINTER(C.c_uint64)]),
        }
        for name, (res, args) in sig.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _LIB = L
    return _LIB


class JavaRandom:
    def __init__(self, seed):
        self.s = C.c_uint64(0)
        lib().o_jr_seed(C.byref(self.s), seed)

    def next_int(self, bound=None):
        return lib().o_jr_next_int(C.byref(self.s)) if bound is None else lib().o_jr_next_int_bound(C.byref(self.s), bound)

    def next_double(self):
        return lib().o_jr_next_double(C.byref(self.s))


class Utt:
    def __init__(self, version=1, conflict=1, handle=None):
        self.h = handle if handle is not None else lib().o_utt_create(version, conflict)
        self.version, self.conflict = version, conflict

    @classmethod
    def from_fields(cls, conflict, types):
        """types: list of (fields[12], flags, produces[])."""
        h = lib().o_utt_empty(conflict)
        for tid, (fields, flags, prod) in enumerate(types):
            f = (C.c_int16 * 12)(*fields)
            p = (C.c_uint8 * max(1, len(prod)))(*prod)
            lib().o_utt_set_type(h, tid, f, flags, len(prod), p)
        return cls(0, conflict, handle=h)

    def field(self, tid, f):
        return lib().o_utt_field(self.h, tid, f)

    def flags(self, tid):
        return lib().o_utt_flags(self.h, tid)

    def produces(self, tid):
        buf = (C.c_uint8 * 16)()
        n = lib().o_utt_produces(self.h, tid, buf)
        return list(buf[:n])

    def max_attack_range(self):
        return lib().o_utt_max_attack_range(self.h)


class Game:
    """One game state (GameState + PhysicalGameState of the reference)."""

    def __init__(self, utt, mapd=None, handle=None):
        self.utt = utt
        if handle is not None:
            self.h = handle
            self.w, self.h_ = mapd
            return
        self.w, self.h_ = mapd["w"], mapd["h"]
        terr = (C.c_uint8 * (self.w * self.h_))(*[int(c) for c in mapd["terrain"]])
        self.h = lib().o_game_create(utt.h, self.w, self.h_, terr, mapd["players"][0][1], mapd["players"][1][1])
        for (tn, uid, pl, x, y, res, hp) in mapd["units"]:
            lib().o_game_add_unit(self.h, TYPE_NAMES.index(tn) if isinstance(tn, str) else tn, uid, pl, x, y, res, hp)

    def __del__(self):
        try:
            lib().o_game_free(self.h)
        except Exception:
            pass

    def clone(self):
        return Game(self.utt, (self.w, self.h_), handle=lib().o_game_clone(self.h))

    def po_view(self, observer):
        return Game(self.utt, (self.w, self.h_), handle=lib().o_po_view(self.h, observer))

    def seed(self, s):
        lib().o_game_seed(self.h, s)

    time = property(lambda self: lib().o_game_time(self.h))
    n_units = property(lambda self: lib().o_game_n_units(self.h))
    winner = property(lambda self: lib().o_game_winner(self.h))
    gameover = property(lambda self: bool(lib().o_game_gameover(self.h)))
    errors = property(lambda self: lib().o_game_errors(self.h))

    def set_rng_state(self, which, state):
        lib().o_game_set_rng_state(self.h, which, state)

    def rng_state(self, which=0):
        return lib().o_game_rng_state(self.h, which)

    def resources(self, p):
        return lib().o_game_resources(self.h, p)

    def units(self):
        """ndarray [n,8]: type, player, x, y, res, hp, id_lo, id_hi (list order)."""
        n = self.n_units
        out = np.zeros((max(n, 1), 8), dtype=np.int32)
        lib().o_game_units(self.h, out.ctypes.data_as(C.POINTER(C.c_int32)))
        return out[:n]

    def assignments(self):
        """ndarray [n,8]: has, type, param, x, y, utype, issue_time, insertion rank (per unit, list order)."""
        n = self.n_units
        out = np.zeros((max(n, 1), 8), dtype=np.int32)
        lib().o_game_assignments(self.h, out.ctypes.data_as(C.POINTER(C.c_int32)))
        return out[:n]

    def cycle(self):
        return bool(lib().o_game_cycle(self.h))

    def is_complete(self):
        return bool(lib().o_game_is_complete(self.h))

    def next_change_time(self):
        return lib().o_game_next_change_time(self.h)

    def issue(self, pairs, safe=True):
        """pairs: list of (unit_list_index, (type,param,x,y,utype))."""
        n = len(pairs)
        idx = (C.c_int32 * max(n, 1))(*[p[0] for p in pairs])
        acts = (ActionV * max(n, 1))(*[ActionV(*p[1]) for p in pairs])
        return lib().o_game_issue(self.h, n, idx, acts, 1 if safe else 0)

    def issue_out(self, pairs, safe=True):
        """issue / issueSafe; returns the PlayerAction as the call left it (issueSafe turns illegal actions into NONE)."""
        n = len(pairs)
        idx = (C.c_int32 * max(n, 1))(*[p[0] for p in pairs])
        acts = (ActionV * max(n, 1))(*[ActionV(*p[1]) for p in pairs])
        lib().o_game_issue_out(self.h, n, idx, acts, 1 if safe else 0)
        return [(idx[k], acts[k].tup()) for k in range(n)]

    def unit_actions(self, unit_idx, none_duration=10):
        buf = (ActionV * 4200)()
        n = lib().o_unit_actions(self.h, unit_idx, none_duration, buf, 4200)
        return [buf[i].tup() for i in range(n)]

    def free_cell(self, x, y):
        return bool(lib().o_game_free_cell(self.h, x, y))

    def _pairs(self, fn, *args):
        cap = self.n_units + 8
        idx = (C.c_int32 * cap)()
        acts = (ActionV * cap)()
        n = fn(*args, idx, acts)
        return [(idx[i], acts[i].tup()) for i in range(n)]

    def random_biased(self, player):
        return self._pairs(lib().o_ai_random_biased, self.h, player)

    def from_vector_action(self, player, vec, fill_none=1):
        vec = np.ascontiguousarray(vec, dtype=np.int32).reshape(-1, 8)
        cap = self.n_units + len(vec) + 8
        idx = (C.c_int32 * cap)()
        acts = (ActionV * cap)()
        n = lib().o_from_vector_action(self.h, player, len(vec), vec.ctypes.data_as(C.POINTER(C.c_int32)),
                                       -9999 if fill_none is None else fill_none, idx, acts)
        return [(idx[i], acts[i].tup()) for i in range(n)]

    def pathfind(self, kind, unit_idx, targetpos, rng, ru=()):
        a = (C.c_int32 * max(1, len(ru)))(*ru)
        return lib().o_pathfind(self.h, kind, unit_idx, targetpos, rng, len(ru), a)

    def observe(self, player, po=False):
        c = 8 if po else 6
        out = np.zeros((c, self.h_, self.w), dtype=np.int32)
        (lib().o_observe_po if po else lib().o_observe)(self.h, player, out.ctypes.data_as(C.POINTER(C.c_int32)))
        return out

    def masks(self, player):
        r = self.utt.max_attack_range() * 2 + 1
        k = 1 + 6 + 16 + 7 + r * r
        out = np.zeros((self.h_, self.w, k), dtype=np.int32)
        lib().o_masks(self.h, player, out.ctypes.data_as(C.POINTER(C.c_int32)))
        return out

    def evaluate(self, fn, maxplayer, minplayer):
        return float(lib().o_evaluate(self.h, fn, maxplayer, minplayer))

    def run_po(self, kind0, ai0, kind1, ai1, n_cycles, max_cycles):
        """Game.start with partiallyObservable = true: each AI decides on its PartiallyObservableGameState view."""
        return bool(lib().o_run_game_po(self.h, kind0, ai0.h if ai0 else None, kind1, ai1.h if ai1 else None, n_cycles, max_cycles))

    def run(self, kind0, ai0, kind1, ai1, n_cycles, max_cycles, stats=None):
        st = (C.c_int64 * 4)() if stats is None else stats
        return bool(lib().o_run_game(self.h, kind0, ai0.h if ai0 else None, kind1, ai1.h if ai1 else None, n_cycles,
                                      max_cycles, st)), list(st)

    def run_observing(self, kind0, kind1, n_cycles, max_cycles):
        """Game.start loop + both players' observations every cycle, entirely in C (bench cpu_baseline); returns cycles run."""
        scratch = np.zeros(6 * self.w * self.h_, dtype=np.int32)
        return lib().o_run_game_observing(self.h, kind0, None, kind1, None, n_cycles, max_cycles, scratch.ctypes.data_as(C.POINTER(C.c_int32)))

    def simulate(self, time_limit):
        return bool(lib().o_simulate(self.h, time_limit))


class ScriptedAI:
    def __init__(self, kind, pathfinder=PF_ASTAR):
        self.kind = kind
        self.h = lib().o_ai_create(kind, pathfinder)

    def __del__(self):
        try:
            lib().o_ai_free(self.h)
        except Exception:
            pass

    def get_action(self, game, player):
        return game._pairs(lib().o_ai_get_action, self.h, game.h, player)


class Pag:
    """PlayerActionGenerator of the oracle (the game must stay alive while the generator is used)."""

    def __init__(self, game, player, none_duration=10):
        self.game = game
        self.h = lib().o_pag_create(game.h, player, none_duration)
        assert self.h, "no unit can act"

    def __del__(self):
        try:
            lib().o_pag_free(self.h)
        except Exception:
            pass

    size = property(lambda self: lib().o_pag_size(self.h))
    generated = property(lambda self: lib().o_pag_generated(self.h))

    def _out(self, fn, *args):
        cap = self.game.n_units + 8
        idx = (C.c_int32 * cap)()
        acts = (ActionV * cap)()
        n = fn(self.h, *args, idx, acts)
        return None if n < 0 else [(idx[k], acts[k].tup()) for k in range(n)]

    def next(self):
        return self._out(lib().o_pag_next)

    def random(self, rng):
        return self._out(lib().o_pag_random, C.byref(rng.s))

    def randomize_order(self, rng):
        lib().o_pag_randomize_order(self.h, C.byref(rng.s))


def player_actions(game, player, max_ints=4000000):
    """GameState.getPlayerActions: (list of [(unit index, action tuple)], total count)."""
    buf = np.zeros(max_ints, dtype=np.int32)
    total = lib().o_player_actions(game.h, player, buf.ctypes.data_as(C.POINTER(C.c_int32)), max_ints)
    out, w = [], 0
    while len(out) < total and w < max_ints and (w + 1 + 6 * int(buf[w])) <= max_ints:
        n = int(buf[w]); w += 1
        out.append([(int(buf[w + 6 * j]), tuple(int(v) for v in buf[w + 6 * j + 1:w + 6 * j + 6])) for j in range(n)])
        w += 6 * n
    return out, total


class Mcts:
    def __init__(self, game, player, seed, lookahead=100, max_depth=10, e_l=0.3, e_g=0.0, e_0=0.4, strategy=0, fensa=True, eval_fn=0):
        self.game = game
        self.h = lib().o_mcts_create(game.h, player, lookahead, max_depth, e_l, e_g, e_0, strategy, 1 if fensa else 0, eval_fn, seed)

    def __del__(self):
        try:
            lib().o_mcts_free(self.h)
        except Exception:
            pass

    def iterate(self, n):
        lib().o_mcts_iterate(self.h, n)

    def root(self, max_children=4096):
        rv, ra = C.c_int32(0), C.c_double(0)
        cv, ca = (C.c_int32 * max_children)(), (C.c_double * max_children)()
        n = lib().o_mcts_root(self.h, C.byref(rv), C.byref(ra), cv, ca, max_children)
        return rv.value, ra.value, np.array(cv[:n], dtype=np.int32), np.array(ca[:n], dtype=np.float64)

    n_nodes = property(lambda self: lib().o_mcts_n_nodes(self.h))

    def best_action(self):
        cap = self.game.n_units + 8
        idx = (C.c_int32 * cap)()
        acts = (ActionV * cap)()
        n = lib().o_mcts_best_action(self.h, idx, acts)
        return None if n < 0 else [(idx[k], acts[k].tup()) for k in range(n)]


class FloodFill:
    """One FloodFillPathFinding instance of the oracle (its distance-map cache persists across find() calls)."""

    def __init__(self):
        self.h = lib().o_ff_create()

    def __del__(self):
        try:
            lib().o_ff_free(self.h)
        except Exception:
            pass

    def find(self, game, unit_idx, targetpos, rng, ru=()):
        a = (C.c_int32 * max(1, len(ru)))(*ru)
        return lib().o_ff_find(self.h, game.h, unit_idx, targetpos, rng, len(ru), a)


class Uct:
    def __init__(self, game, player, seed, lookahead=100, max_depth=10, eval_fn=0):
        self.game = game
        self.h = lib().o_uct_create(game.h, player, lookahead, max_depth, eval_fn, seed)

    def __del__(self):
        try:
            lib().o_uct_free(self.h)
        except Exception:
            pass

    def iterate(self, n):
        lib().o_uct_iterate(self.h, n)

    def root(self, max_children=4096):
        rv, ra = C.c_int32(0), C.c_float(0)
        cv, ca = (C.c_int32 * max_children)(), (C.c_float * max_children)()
        n = lib().o_uct_root(self.h, C.byref(rv), C.byref(ra), cv, ca, max_children)
        return rv.value, float(ra.value), np.array(cv[:n], dtype=np.int32), np.array(ca[:n], dtype=np.float64)

    n_nodes = property(lambda self: lib().o_uct_n_nodes(self.h))

    def best_action(self):
        cap = self.game.n_units + 8
        idx = (C.c_int32 * cap)()
        acts = (ActionV * cap)()
        n = lib().o_uct_best_action(self.h, idx, acts)
        return None if n < 0 else [(idx[k], acts[k].tup()) for k in range(n)]
