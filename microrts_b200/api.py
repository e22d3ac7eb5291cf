"""Host-side mirror of the reference interface for the batched simulation path.

  UnitTypeTable       <- src/rts/units/UnitTypeTable.java:92-94,104-289
  PhysicalGameState   <- src/rts/PhysicalGameState.java:65-76,700-726 (load / fromXML)
  UnitAction          <- src/rts/UnitAction.java:29-100 (constants)
  BatchedGameState    <- the reference's GameState API (issueSafe/issue/cycle/winner/gameover/getVectorObservation,
                         src/rts/GameState.java) applied to n games at once, plus the JNIGridnetVecClient-style
                         reset/gameStep/getMasks surface (src/tests/JNIGridnetVecClient.java:179-316)

All arrays are numpy (host) unless a torch CUDA tensor is passed as `out`/input, in which case its device pointer is
handed to the C ABI directly (no copies).
"""
import ctypes as C

import numpy as np

from . import _ffi

ACTIONS_VECTOR, ACTIONS_RAW = 0, 1
(POLICY_EXTERNAL, POLICY_PASSIVE, POLICY_RANDOM_BIASED, POLICY_WORKER_RUSH, POLICY_LIGHT_RUSH, POLICY_HEAVY_RUSH, POLICY_RANGED_RUSH,
 POLICY_WORKER_DEFENSE, POLICY_LIGHT_DEFENSE, POLICY_HEAVY_DEFENSE, POLICY_RANGED_DEFENSE,
 POLICY_PO_WORKER_RUSH, POLICY_PO_LIGHT_RUSH, POLICY_PO_HEAVY_RUSH, POLICY_PO_RANGED_RUSH, POLICY_WORKER_RUSH_PP, POLICY_CRUSH_V1, POLICY_CRUSH_V2, POLICY_EMR_DETERMINISTICO) = range(19)
PF_ASTAR, PF_BFS, PF_GREEDY, PF_FLOODFILL = 0, 1, 2, 3
DTYPE_U8, DTYPE_I32, DTYPE_BITS = 0, 1, 2
FLAG_PARTIAL_OBS = 1
FLAG_SCRIPTED_AI = 2
FLAG_PO_POLICIES = 4


class MicroRTSError(RuntimeError):
    pass


def _check(rc):
    if rc < 0:
        raise MicroRTSError("libmicrorts_cuda error %d: %s" % (rc, _ffi.lib().mrts_last_error().decode()))
    return rc


class UnitAction:
    TYPE_NONE, TYPE_MOVE, TYPE_HARVEST, TYPE_RETURN, TYPE_PRODUCE, TYPE_ATTACK_LOCATION = range(6)
    NUMBER_OF_ACTION_TYPES = 6
    DIRECTION_NONE, DIRECTION_UP, DIRECTION_RIGHT, DIRECTION_DOWN, DIRECTION_LEFT = -1, 0, 1, 2, 3
    DIRECTION_OFFSET_X = (0, 1, 0, -1)
    DIRECTION_OFFSET_Y = (-1, 0, 1, 0)


class UnitType:
    FIELDS = ["cost", "hp", "minDamage", "maxDamage", "attackRange", "produceTime", "moveTime", "attackTime",
              "harvestTime", "returnTime", "harvestAmount", "sightRadius"]

    def __init__(self, utt, tid):
        L = _ffi.lib()
        self.ID = tid
        self.name = L.mrts_utt_type_name(utt._h, tid).decode()
        for k, f in enumerate(self.FIELDS):
            setattr(self, f, L.mrts_utt_get(utt._h, tid, k))
        fl = L.mrts_utt_get(utt._h, tid, 12)
        self.isResource, self.isStockpile, self.canHarvest, self.canMove, self.canAttack = [bool(fl >> b & 1) for b in range(5)]
        n = L.mrts_utt_get(utt._h, tid, 13)
        self.produces = [L.mrts_utt_get(utt._h, tid, 14 + k) for k in range(n)]


class UnitTypeTable:
    VERSION_ORIGINAL, VERSION_ORIGINAL_FINETUNED, VERSION_NON_DETERMINISTIC = 1, 2, 3
    MOVE_CONFLICT_RESOLUTION_CANCEL_BOTH, MOVE_CONFLICT_RESOLUTION_CANCEL_RANDOM, MOVE_CONFLICT_RESOLUTION_CANCEL_ALTERNATING = 1, 2, 3

    def __init__(self, version=1, conflict_policy=1, _handle=None):
        if _handle is None:
            h = C.c_void_p()
            _check(_ffi.lib().mrts_utt_create(version, conflict_policy, C.byref(h)))
            _handle = h
        self._h = _handle
        self._types = [UnitType(self, t) for t in range(_ffi.lib().mrts_utt_num_types(self._h))]

    @classmethod
    def fromJSON(cls, text):
        h = C.c_void_p()
        _check(_ffi.lib().mrts_utt_from_json(text.encode(), C.byref(h)))
        return cls(_handle=h)

    def __del__(self):
        try:
            _ffi.lib().mrts_utt_destroy(self._h)
        except Exception:
            pass

    def getUnitTypes(self):
        return list(self._types)

    def getUnitType(self, key):
        if isinstance(key, str):
            for t in self._types:
                if t.name == key:
                    return t
            return None
        return self._types[key]

    def getMoveConflictResolutionStrategy(self):
        return _ffi.lib().mrts_utt_conflict_policy(self._h)

    def getMaxAttackRange(self):
        return _ffi.lib().mrts_utt_max_attack_range(self._h)


class PhysicalGameState:
    """A map / initial state.  Built by load(path, utt), fromXML(text, utt) or create(...)."""

    def __init__(self, handle, utt):
        self._h, self.utt = handle, utt

    @classmethod
    def load(cls, fileName, utt):
        h = C.c_void_p()
        _check(_ffi.lib().mrts_map_load_xml(str(fileName).encode(), utt._h, C.byref(h)))
        return cls(h, utt)

    @classmethod
    def fromXML(cls, text, utt):
        h = C.c_void_p()
        _check(_ffi.lib().mrts_map_from_xml(text.encode(), utt._h, C.byref(h)))
        return cls(h, utt)

    @classmethod
    def create(cls, width, height, terrain, resources, units, utt):
        """units: rows [type, id, player, x, y, resources, hitpoints]."""
        t = np.ascontiguousarray(terrain, dtype=np.uint8).reshape(-1)
        u = np.ascontiguousarray(units, dtype=np.int32).reshape(-1, 7)
        h = C.c_void_p()
        _check(_ffi.lib().mrts_map_create(width, height, t.ctypes.data, resources[0], resources[1], len(u),
                                          u.ctypes.data if len(u) else None, utt._h, C.byref(h)))
        return cls(h, utt)

    def __del__(self):
        try:
            _ffi.lib().mrts_map_destroy(self._h)
        except Exception:
            pass

    def getWidth(self):
        return _ffi.lib().mrts_map_width(self._h)

    def getHeight(self):
        return _ffi.lib().mrts_map_height(self._h)

    def getUnits(self):
        n = _ffi.lib().mrts_map_num_units(self._h)
        out = np.zeros((max(n, 1), 7), dtype=np.int32)
        _ffi.lib().mrts_map_get_units(self._h, out.ctypes.data)
        return out[:n]

    def getTerrain(self):
        out = np.zeros(self.getWidth() * self.getHeight(), dtype=np.uint8)
        _ffi.lib().mrts_map_get_terrain(self._h, out.ctypes.data)
        return out.reshape(self.getHeight(), self.getWidth())

    def getPlayerResources(self, p):
        return _ffi.lib().mrts_map_resources(self._h, p)


def _ptr(a):
    """(pointer, on_device, keepalive) of a numpy array or torch tensor."""
    if a is None:
        return None, 0, None
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data, 0, a
    # torch tensor
    assert a.is_contiguous()
    return a.data_ptr(), 1 if a.is_cuda else 0, a


class BatchedGameState:
    """n independent GameState objects stepped in lockstep on one GPU (rts.cuda.BatchedGameState)."""

    def __init__(self, utt, pgs, n_games, device=0, partial_obs=False, unit_capacity=0, scripted_ai=False, po_policies=False):
        """po_policies: Game(partiallyObservable=true) -- device policies decide on their player's PartiallyObservableGameState view."""
        maps = list(pgs) if isinstance(pgs, (list, tuple)) else [pgs]
        self.utt, self.maps, self.n = utt, maps, int(n_games)
        arr = (C.c_void_p * len(maps))(*[m._h for m in maps])
        h = C.c_void_p()
        _check(_ffi.lib().mrts_batch_create(utt._h, arr, len(maps), self.n, device,
                                            (FLAG_PARTIAL_OBS if partial_obs else 0) | (FLAG_SCRIPTED_AI if scripted_ai else 0) | (FLAG_PO_POLICIES if po_policies else 0),
                                            unit_capacity, C.byref(h)))
        self._h = h
        self.width, self.height = maps[0].getWidth(), maps[0].getHeight()
        self.cap = _ffi.lib().mrts_batch_unit_capacity(h)
        self.device = device
        self.num_planes = _ffi.lib().mrts_batch_num_planes(h)
        self.mask_width = _ffi.lib().mrts_batch_mask_width(h)

    def close(self):
        if self._h is not None:
            _ffi.lib().mrts_batch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- lifecycle ---------------------------------------------------------------------------------------------------
    def reset(self, seeds=None):
        p, dev, _k = _ptr(None if seeds is None else (np.ascontiguousarray(seeds, dtype=np.int64) if isinstance(seeds, (list, np.ndarray)) else seeds))
        _check(_ffi.lib().mrts_batch_reset(self._h, p, dev))

    def reset_masked(self, mask, seeds=None):
        m = np.ascontiguousarray(mask, dtype=np.uint8) if isinstance(mask, (list, np.ndarray)) else mask
        s = None if seeds is None else (np.ascontiguousarray(seeds, dtype=np.int64) if isinstance(seeds, (list, np.ndarray)) else seeds)
        pm, dm, _k1 = _ptr(m)
        ps, ds, _k2 = _ptr(s)
        assert s is None or dm == ds
        _check(_ffi.lib().mrts_batch_reset_masked(self._h, pm, ps, dm))

    def copy_games(self, src, src_index=None, mask=None):
        """GameState.clone() batched: game g becomes a copy of src's game src_index[g] (default g) where mask[g] (default all)."""
        idx = None if src_index is None else (np.ascontiguousarray(src_index, dtype=np.int64) if isinstance(src_index, (list, np.ndarray)) else src_index)
        m = None if mask is None else (np.ascontiguousarray(mask, dtype=np.uint8) if isinstance(mask, (list, np.ndarray)) else mask)
        pi, di, _k1 = _ptr(idx)
        pm, dm, _k2 = _ptr(m)
        assert idx is None or m is None or di == dm
        _check(_ffi.lib().mrts_batch_copy_games(self._h, src._h, pi, pm, di if idx is not None else dm))

    def set_policy(self, player, policy, pathfinder=PF_ASTAR):
        _check(_ffi.lib().mrts_batch_set_policy(self._h, player, policy, pathfinder))

    def restart_masked(self, mask):
        """Restart the masked games from their map; their RNG streams keep running (JNIGridnetVecClient auto-reset)."""
        m = np.ascontiguousarray(mask, dtype=np.uint8) if isinstance(mask, (list, np.ndarray)) else mask
        pm, dm, _k = _ptr(m)
        _check(_ffi.lib().mrts_batch_restart_masked(self._h, pm, dm))

    def set_issue_order(self, sequential):
        """False: both players decide on the pre-issue state (Game.start); True: player 1 decides after player 0's actions
        are issued (JNIGridnetClientSelfPlay.gameStep)."""
        _check(_ffi.lib().mrts_batch_set_issue_order(self._h, 1 if sequential else 0))

    def set_info_output(self, out):
        """Device buffer int32 [n][2][12] that every later step() fills with the reward functions' step facts (None: off)."""
        if out is not None:
            assert tuple(out.shape) == (self.n, 2, 12) and "int32" in str(out.dtype)
        self._info_keepalive = out
        _check(_ffi.lib().mrts_batch_set_info_output(self._h, _ptr(out)[0] if out is not None else None))

    def set_auto_reset(self, enable=True):
        _check(_ffi.lib().mrts_batch_set_auto_reset(self._h, 1 if enable else 0))

    def sync(self):
        _check(_ffi.lib().mrts_batch_sync(self._h))

    # -- actions -----------------------------------------------------------------------------------------------------
    @staticmethod
    def _actions(actions, counts):
        if isinstance(actions, (list, np.ndarray)):
            actions = np.ascontiguousarray(actions, dtype=np.int32)
        if isinstance(counts, (list, np.ndarray)):
            counts = np.ascontiguousarray(counts, dtype=np.int32)
        max_k = int(actions.shape[1]) if actions is not None and len(actions.shape) == 3 else 0
        return actions, counts, max_k

    def set_actions(self, player, actions, counts=None, fmt=ACTIONS_VECTOR, fill_none_duration=1):
        """Stage one PlayerAction per game ([n][max_k][8] rows) for the next step() of an EXTERNAL player."""
        a, c, k = self._actions(actions, counts)
        pa, da, _k1 = _ptr(a)
        pc, dc, _k2 = _ptr(c)
        assert c is None or a is None or dc == da, "actions and counts must both be host arrays or both device tensors"
        _check(_ffi.lib().mrts_batch_set_actions(self._h, player, fmt, pa, pc, k, fill_none_duration, da))

    def issue(self, player, actions, counts=None, fmt=ACTIONS_RAW, safe=False, fill_none_duration=-1):
        """GameState.issue(pa) (GameState.java:249-328) for every game; issueSafe when safe=True."""
        a, c, k = self._actions(actions, counts)
        pa, da, _k1 = _ptr(a)
        pc, dc, _k2 = _ptr(c)
        assert c is None or a is None or dc == da, "actions and counts must both be host arrays or both device tensors"
        _check(_ffi.lib().mrts_batch_issue(self._h, player, fmt, pa, pc, k, fill_none_duration, 1 if safe else 0, da))

    def issueSafe(self, player, actions, counts=None, fmt=ACTIONS_RAW, fill_none_duration=-1):
        """GameState.issueSafe(pa) (GameState.java:338-408)."""
        self.issue(player, actions, counts, fmt, True, fill_none_duration)

    # -- time --------------------------------------------------------------------------------------------------------
    def step(self, n_cycles=1, max_cycles=3000):
        """Game.start loop body (Game.java:126-140) for n_cycles cycles per game."""
        _check(_ffi.lib().mrts_batch_step(self._h, n_cycles, max_cycles))

    def cycle(self, n_cycles=1):
        """GameState.cycle() n times (no policies)."""
        _check(_ffi.lib().mrts_batch_cycle_to(self._h, None, n_cycles, 0))

    def cycle_to(self, t_target):
        t = np.ascontiguousarray(t_target, dtype=np.int32) if isinstance(t_target, (list, np.ndarray)) else t_target
        p, dev, _k = _ptr(t)
        _check(_ffi.lib().mrts_batch_cycle_to(self._h, p, 0, dev))

    def cycle_to_decision(self):
        """cycle() until the game is over or some player has a unit without an assignment (the loop at the head of an MCTS node)."""
        _check(_ffi.lib().mrts_batch_cycle_to_decision(self._h))

    def unit_actions(self, player, none_duration=10, max_choices=None, max_actions=64):
        """Unit.getUnitActions of every idle unit of `player`, ordered (mrts_batch_unit_actions).  Returns per game a dict
        {time, gameover, winner, can_act: (p0, p1), resources, resources_used, positions_used, choices: [(slot, id, type, x, y, [actions])]}
        with actions as (type, parameter, x, y, unit type) tuples like the oracle's (parameter = direction, or the duration of NONE)."""
        K = max_choices or self.cap
        hdr = np.zeros((self.n, 8), dtype=np.int32)
        pos = np.zeros((self.n, self.cap), dtype=np.int32)
        ch = np.zeros((self.n, K, 4), dtype=np.int32)
        ls = np.zeros((self.n, K, max_actions), dtype=np.int32)
        _check(_ffi.lib().mrts_batch_unit_actions(self._h, player, none_duration, K, max_actions, hdr.ctypes.data, pos.ctypes.data, ch.ctypes.data, ls.ctypes.data, 0))
        out = []
        for g in range(self.n):
            h = hdr[g]
            choices = []
            for c in range(min(int(h[0]), K)):
                slot, uid, packed, cnt = [int(v) for v in ch[g, c]]
                acts = []
                for k in range(min(cnt, max_actions)):
                    v = int(ls[g, c, k]) & 0xffffffff
                    at, d, x, y, ut = v & 15, ((v >> 4) & 15) - 1, (v >> 8) & 255, (v >> 16) & 255, ((v >> 24) & 255) - 1
                    acts.append((at, none_duration if at == 0 else d, x, y, ut))
                choices.append((slot, uid, packed & 255, (packed >> 8) & 255, (packed >> 16) & 255, acts))
            out.append(dict(time=int(h[6]), gameover=bool(h[7] & 1), winner=((int(h[7]) >> 1) & 3) - 1, can_act=(bool(h[7] & 8), bool(h[7] & 16)),
                            resources=(int(h[3]), int(h[4])), resources_used=(int(h[1]), int(h[2])), positions_used=[int(v) for v in pos[g, :int(h[5])]],
                            choices=choices))
        return out

    def evaluate(self, eval_fn=0, maxplayer=0, observer=-1):
        """EvaluationFunction.evaluate(maxplayer, 1 - maxplayer, gs) of every game's current state (0 = SimpleSqrtEvaluationFunction3,
        1 = SimpleEvaluationFunction); observer >= 0: of that player's partially observable view."""
        out = np.empty(self.n, dtype=np.float32)
        _check(_ffi.lib().mrts_batch_evaluate(self._h, eval_fn, maxplayer, observer, out.ctypes.data, 0))
        return out

    def find_path(self, pathfinder, unit_cells, target_positions, ranges):
        """PathFinding.findPathToPositionInRange for one unit per game: the unit on cell unit_cells[g] (x + y*W) towards
        target_positions[g] within ranges[g] (range < 0: findPath).  Returns the MOVE direction per game, -1 for null."""
        q = np.ascontiguousarray(np.stack([np.broadcast_to(np.asarray(a, dtype=np.int32), (self.n,)) for a in (unit_cells, target_positions, ranges)], axis=1))
        out = np.empty(self.n, dtype=np.int32)
        _check(_ffi.lib().mrts_batch_pathfind(self._h, pathfinder, q.ctypes.data, out.ctypes.data, 0))
        return out

    def rollout(self, depth=100, rollouts_per_game=1, eval_fn=0, maxplayer=0, observer=-1, seeds=None):
        """NaiveMCTS.simulate + evaluation for every game (the batch is not modified).
        Returns (evaluation[float32], simulated_cycles[int32]) of shape [n, rollouts_per_game]; the reference's discounted
        value is evaluation * 0.99 ** (cycles / 10.0) (NaiveMCTS.java:205)."""
        nr = self.n * rollouts_per_game
        ev = np.zeros(nr, dtype=np.float32)
        tm = np.zeros(nr, dtype=np.int32)
        s = None if seeds is None else np.ascontiguousarray(seeds, dtype=np.int64).reshape(-1)
        assert s is None or len(s) == nr
        _check(_ffi.lib().mrts_batch_rollout(self._h, rollouts_per_game, depth, eval_fn, maxplayer, observer,
                                             None if s is None else s.ctypes.data, ev.ctypes.data, tm.ctypes.data, 0))
        return ev.reshape(self.n, rollouts_per_game), tm.reshape(self.n, rollouts_per_game)

    # -- outputs -----------------------------------------------------------------------------------------------------
    def observe(self, player, dtype=np.int32, out=None):
        """GameState.getVectorObservation(player) for every game: [n][C][H][W]."""
        code = DTYPE_U8 if np.dtype(dtype) == np.uint8 else DTYPE_I32
        if out is None:
            out = np.empty((self.n, self.num_planes, self.height, self.width), dtype=np.uint8 if code == DTYPE_U8 else np.int32)
        p, dev, _k = _ptr(out)
        _check(_ffi.lib().mrts_batch_observe(self._h, player, code, p, dev))
        return out

    def set_observation_outputs(self, out0=None, out1=None):
        """Fused emission: every later step() also writes getVectorObservation(0) / (1) of the state it leaves behind into
        out0 / out1 ([n][6][H][W], uint8 or int32 device tensors of the batch's device; None disables a player)."""
        outs = [o for o in (out0, out1) if o is not None]
        code = DTYPE_I32
        if outs:
            dt = str(outs[0].dtype)
            code = DTYPE_U8 if "uint8" in dt else DTYPE_I32
            for o in outs:
                assert str(o.dtype) == dt and tuple(o.shape) == (self.n, self.num_planes, self.height, self.width)
        ptrs = [_ptr(o)[0] if o is not None else None for o in (out0, out1)]
        self._obs_keepalive = (out0, out1)
        _check(_ffi.lib().mrts_batch_set_observation_outputs(self._h, code, ptrs[0], ptrs[1]))

    def set_output_layout(self, obs, masks=None, interleaved=False, side=0):
        """Fused outputs in environment order (JNIGridnetVecClient): obs = [E][C][H][W] and masks = [E][H][W][(mask_width + 7) // 8]
        device arrays.  interleaved: E = 2n, environment 2g is player 0 of game g and 2g + 1 player 1 (self-play pairs); otherwise
        E = n and the arrays hold player `side` only."""
        code = DTYPE_U8 if "uint8" in str(obs.dtype) else DTYPE_I32
        E = 2 * self.n if interleaved else self.n
        assert tuple(obs.shape) == (E, self.num_planes, self.height, self.width)
        opg = self.num_planes * self.height * self.width * (1 if code == DTYPE_U8 else 4)
        base = _ptr(obs)[0]
        ptrs = [base, base + opg] if interleaved else [base if side == 0 else None, base if side == 1 else None]
        mptrs = [None, None]
        if masks is not None:
            mb = (self.mask_width + 7) // 8
            assert tuple(masks.shape) == (E, self.height, self.width, mb) and "uint8" in str(masks.dtype)
            mbase = _ptr(masks)[0]
            mptrs = [mbase, mbase + self.height * self.width * mb] if interleaved else [mbase if side == 0 else None, mbase if side == 1 else None]
        self._obs_keepalive, self._mask_keepalive = obs, masks
        L = _ffi.lib()
        _check(L.mrts_batch_set_output_stride(self._h, 2 if interleaved else 1))
        _check(L.mrts_batch_set_observation_outputs(self._h, code, ptrs[0], ptrs[1]))
        _check(L.mrts_batch_set_mask_outputs(self._h, mptrs[0], mptrs[1]))

    def set_vec_autoreset(self, done_mode, max_steps):
        """JNIGridnetVecClient's auto-reset inside the step launch (mrts_batch_set_vec_autoreset)."""
        _check(_ffi.lib().mrts_batch_set_vec_autoreset(self._h, done_mode, max_steps))

    def set_actions_interleaved(self, actions, fmt=ACTIONS_VECTOR, fill_none_duration=1, async_copy=False):
        """Both players' rows of every game from one [2n][max_k][8] array in self-play environment order (2g = player 0 of game g)."""
        if isinstance(actions, (list, np.ndarray)):
            actions = np.ascontiguousarray(actions, dtype=np.int32)
        assert actions.shape[0] == 2 * self.n and actions.shape[2] == 8
        pa, da, _k = _ptr(actions)
        _check(_ffi.lib().mrts_batch_set_actions_interleaved(self._h, fmt, pa, int(actions.shape[1]), fill_none_duration, da, 1 if async_copy else 0))

    def vec_step(self, actions, async_copy=True):
        """set_actions_interleaved + step(1) in one library call (JNIGridnetVecClient.gameStep of self-play environments)."""
        pa, da, _k = _ptr(actions)
        _check(_ffi.lib().mrts_batch_vec_step(self._h, pa, int(actions.shape[1]), da, 1 if async_copy else 0))

    def set_mask_outputs(self, out0=None, out1=None):
        """Fused emission of the bit-packed action masks: every later step() also writes getMasks(0) / (1) of the state it leaves
        behind into out0 / out1 ([n][H][W][(mask_width + 7) // 8] uint8 device tensors; None disables a player)."""
        for o in (out0, out1):
            assert o is None or (tuple(o.shape) == (self.n, self.height, self.width, (self.mask_width + 7) // 8) and "uint8" in str(o.dtype))
        self._mask_keepalive = (out0, out1)
        _check(_ffi.lib().mrts_batch_set_mask_outputs(self._h, _ptr(out0)[0] if out0 is not None else None, _ptr(out1)[0] if out1 is not None else None))

    def masks(self, player, dtype=np.int32, out=None):
        """JNIGridnetClient.getMasks(player) for every game: [n][H][W][mask_width]; dtype="bits": [n][H][W][(mask_width+7)//8]
        uint8 with element j in bit j & 7 of byte j >> 3 (np.unpackbits(..., bitorder="little") restores the dense form)."""
        if isinstance(dtype, str) and dtype == "bits":
            code = DTYPE_BITS
            if out is None:
                out = np.empty((self.n, self.height, self.width, (self.mask_width + 7) // 8), dtype=np.uint8)
        else:
            code = DTYPE_U8 if np.dtype(dtype) == np.uint8 else DTYPE_I32
        if out is None:
            out = np.empty((self.n, self.height, self.width, self.mask_width), dtype=np.uint8 if code == DTYPE_U8 else np.int32)
        p, dev, _k = _ptr(out)
        _check(_ffi.lib().mrts_batch_masks(self._h, player, code, p, dev))
        return out

    def results(self, out=None):
        """[n][4] = time, winner (-1 none), gameover, error bits."""
        if out is None:
            out = np.empty((self.n, 4), dtype=np.int32)
        p, dev, _k = _ptr(out)
        _check(_ffi.lib().mrts_batch_results(self._h, p, dev))
        return out

    def copy_to_host(self, host, dev):
        """Queue an asynchronous device -> host copy of `dev` (device tensor) into `host` (numpy array, ideally pinned) on the batch's
        stream; sync() waits for it."""
        assert host.nbytes == dev.numel() * dev.element_size() if hasattr(dev, "numel") else host.nbytes == dev.nbytes
        _check(_ffi.lib().mrts_batch_copy_to_host(self._h, host.ctypes.data, _ptr(dev)[0], host.nbytes))

    def stats(self):
        out = np.zeros(8, dtype=np.int64)
        _check(_ffi.lib().mrts_batch_stats(self._h, out.ctypes.data))
        return dict(zip(["wins_p0", "wins_p1", "draws", "games_finished", "cycles", "decisions", "unit_cycles", "errors"], out.tolist()))

    def io_bytes(self):
        """(bytes read, bytes written) of game state and outputs by the batch's kernels since the last full reset."""
        out = np.zeros(2, dtype=np.int64)
        _check(_ffi.lib().mrts_batch_io_bytes(self._h, out.ctypes.data))
        return int(out[0]), int(out[1])

    @property
    def last_kernel(self):
        return _ffi.lib().mrts_batch_last_kernel(self._h).decode()

    @property
    def launch_count(self):
        return _ffi.lib().mrts_batch_launch_count(self._h)

    def export(self, first=0, count=None):
        count = self.n - first if count is None else count
        hdr = np.zeros((count, 8), dtype=np.int32)
        units = np.zeros((count, self.cap, 8), dtype=np.int32)
        acts = np.zeros((count, self.cap, 8), dtype=np.int32)
        rng = np.zeros((count, 3), dtype=np.int64)
        st = _ffi.StateHost(hdr.ctypes.data_as(C.POINTER(C.c_int32)), units.ctypes.data_as(C.POINTER(C.c_int32)),
                            acts.ctypes.data_as(C.POINTER(C.c_int32)), rng.ctypes.data_as(C.POINTER(C.c_int64)))
        _check(_ffi.lib().mrts_batch_export(self._h, first, count, C.byref(st)))
        return dict(header=hdr, units=units, actions=acts, rng=rng)

    def import_(self, state, first=0):
        hdr = np.ascontiguousarray(state["header"], dtype=np.int32)
        units = np.ascontiguousarray(state["units"], dtype=np.int32)
        acts = np.ascontiguousarray(state["actions"], dtype=np.int32) if state.get("actions") is not None else None
        rng = np.ascontiguousarray(state["rng"], dtype=np.int64) if state.get("rng") is not None else None
        st = _ffi.StateHost(hdr.ctypes.data_as(C.POINTER(C.c_int32)), units.ctypes.data_as(C.POINTER(C.c_int32)),
                            acts.ctypes.data_as(C.POINTER(C.c_int32)) if acts is not None else None,
                            rng.ctypes.data_as(C.POINTER(C.c_int64)) if rng is not None else None)
        _check(_ffi.lib().mrts_batch_import(self._h, first, len(hdr), C.byref(st)))

    # reference-style accessors over an export
    def getTime(self):
        return self.results()[:, 0]

    def winner(self):
        return self.results()[:, 1]

    def gameover(self):
        return self.results()[:, 2].astype(bool)
