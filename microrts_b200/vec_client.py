"""JNIGridnetVecClient over the batched engine -- the reference's vectorised RL facade (src/tests/JNIGridnetVecClient.java)
with the same constructor arguments, methods and return conventions, backed by libmicrorts_cuda.so instead of one Java
GameState per environment.

  JNIGridnetVecClient(num_selfplay_envs, num_envs, max_steps, rfs, micrortsPath, mapPaths, ai2s, utt, partial_obs)  :106
  reset(players) -> Responses(observation, reward, done)                                                          :179
  gameStep(action, players) -> Responses                                                                          :213
  getMasks(player) -> int[envs][H][W][mask width]                                                                 :307
  close()                                                                                                         :318

Environment layout follows the reference: the first num_selfplay_envs entries are pairs (player 0 and player 1 of the same
game, JNIGridnetClientSelfPlay), the next num_envs entries are agent-vs-bot games (JNIGridnetClient).  What differs:

* ai2s are device policies, named like the reference's AI classes (microrts_b200.ai: PassiveAI, RandomBiasedAI, WorkerRush,
  LightRush); all maps of one client must have the same size (one batch per group of environments).
* The per-cycle order inside an environment is the reference's: self-play pairs decide sequentially (player 1's
  PlayerAction is built on the state that already holds player 0's, JNIGridnetClientSelfPlay.java:160-170); agent-vs-bot
  games build both PlayerActions on the pre-issue state (JNIGridnetClient.java:168-179).
* Rewards are the reference's reward functions (microrts_b200.rewards) evaluated from per-game step facts the step kernel
  writes (no trace objects); auto-reset keeps the terminal reward/done and forces done[0] (:272-286).
* Partially observable observations are taken from the state after the cycle (the reference returns a view built before
  the cycle, JNIGridnetClient.java:163-203) -- DESIGN.md, known deviations.
* One kernel launch per gameStep and group: decode of the vector actions, issueSafe of both players, cycle, reward facts, the
  auto-reset of finished environments (mrts_batch_set_vec_autoreset) and the observations (optionally the bit-packed masks) of every
  environment, written in environment order.  Host arrays are pinned and reused; `compact=True` returns uint8 observations and keeps
  the masks bit-packed (getMasksPacked) -- 4 KB instead of 87 KB per 16x16 environment and step across PCIe.
"""
import os

import numpy as np

from . import api as M

NO_CAP = 1 << 30


class Responses:
    """tests.jni.Responses (observation, reward, done)."""

    def __init__(self, observation, reward, done):
        self.observation, self.reward, self.done = observation, reward, done

    def set(self, observation, reward, done):
        self.observation, self.reward, self.done = observation, reward, done


class AISpec:
    def __init__(self, policy, pathfinder=M.PF_ASTAR):
        self.policy, self.pathfinder = policy, pathfinder

    def key(self):
        return (self.policy, self.pathfinder)


class ai:
    """Device-resident stand-ins for the reference's AI classes accepted as `ai2s` (names as in src/ai)."""

    @staticmethod
    def PassiveAI(utt=None):
        return AISpec(M.POLICY_PASSIVE)

    @staticmethod
    def RandomBiasedAI(utt=None):
        return AISpec(M.POLICY_RANDOM_BIASED)

    @staticmethod
    def WorkerRush(utt=None, pathfinder=M.PF_ASTAR):
        return AISpec(M.POLICY_WORKER_RUSH, pathfinder)

    @staticmethod
    def LightRush(utt=None, pathfinder=M.PF_ASTAR):
        return AISpec(M.POLICY_LIGHT_RUSH, pathfinder)

    @staticmethod
    def HeavyRush(utt=None, pathfinder=M.PF_ASTAR):
        return AISpec(M.POLICY_HEAVY_RUSH, pathfinder)

    @staticmethod
    def RangedRush(utt=None, pathfinder=M.PF_ASTAR):
        return AISpec(M.POLICY_RANGED_RUSH, pathfinder)

    @staticmethod
    def POWorkerRush(utt=None, pathfinder=M.PF_ASTAR):
        return AISpec(M.POLICY_PO_WORKER_RUSH, pathfinder)

    @staticmethod
    def POLightRush(utt=None, pathfinder=M.PF_ASTAR):
        return AISpec(M.POLICY_PO_LIGHT_RUSH, pathfinder)

    @staticmethod
    def POHeavyRush(utt=None, pathfinder=M.PF_ASTAR):
        return AISpec(M.POLICY_PO_HEAVY_RUSH, pathfinder)

    @staticmethod
    def PORangedRush(utt=None, pathfinder=M.PF_ASTAR):
        return AISpec(M.POLICY_PO_RANGED_RUSH, pathfinder)

    @staticmethod
    def WorkerRushPlusPlus(utt=None, pathfinder=M.PF_ASTAR):
        return AISpec(M.POLICY_WORKER_RUSH_PP, pathfinder)

    @staticmethod
    def CRush_V1(utt=None, pathfinder=M.PF_ASTAR):
        return AISpec(M.POLICY_CRUSH_V1, pathfinder)

    @staticmethod
    def CRush_V2(utt=None, pathfinder=M.PF_ASTAR):
        return AISpec(M.POLICY_CRUSH_V2, pathfinder)

    @staticmethod
    def WorkerDefense(utt=None, pathfinder=M.PF_ASTAR):
        return AISpec(M.POLICY_WORKER_DEFENSE, pathfinder)

    @staticmethod
    def LightDefense(utt=None, pathfinder=M.PF_ASTAR):
        return AISpec(M.POLICY_LIGHT_DEFENSE, pathfinder)

    @staticmethod
    def HeavyDefense(utt=None, pathfinder=M.PF_ASTAR):
        return AISpec(M.POLICY_HEAVY_DEFENSE, pathfinder)

    @staticmethod
    def RangedDefense(utt=None, pathfinder=M.PF_ASTAR):
        return AISpec(M.POLICY_RANGED_DEFENSE, pathfinder)


def _device_array(shape, dtype, emulated):
    """Memory the engine writes 'on device': a torch CUDA tensor, or host memory when the library is the test emulator."""
    if emulated:
        return np.zeros(shape, dtype=dtype)
    import torch
    return torch.zeros(shape, dtype=torch.int32 if dtype == np.int32 else torch.uint8, device="cuda")


def _host_array(shape, dtype, emulated):
    """Pinned host memory (a numpy view of a pinned torch tensor) that device results are copied into asynchronously."""
    if emulated:
        return np.zeros(shape, dtype=dtype), None
    import torch
    t = torch.zeros(shape, dtype={np.int32: torch.int32, np.uint8: torch.uint8}[dtype]).pin_memory()
    return t.numpy(), t


def _host(a):
    return a if isinstance(a, np.ndarray) else a.cpu().numpy()


class _Group:
    """One batch: either the self-play games (both players EXTERNAL, sequential issue) or the agent-vs-bot games that share
    one opponent policy and one agent side.  Its environments are envs[0], envs[1], ... of the client; self-play environments come
    in pairs (player 0, player 1 of one game) and the group's output arrays are laid out in that environment order."""

    def __init__(self, utt, maps, envs, selfplay, spec, side, partial_obs, device, emulated, seed0, rfs, max_steps, compact, fused_masks):
        self.envs, self.selfplay, self.side, self.emulated = np.asarray(envs), selfplay, side, emulated
        n = len(envs) // 2 if selfplay else len(envs)
        self.n, self.E = n, len(envs)
        same = all(m is maps[0] for m in maps)
        scripted = spec is not None and spec.policy >= M.POLICY_WORKER_RUSH
        # partial_obs: the opponent decides on ITS PartiallyObservableGameState, as ai2.getAction(1 - player, player2gs) does
        # (JNIGridnetClient.java:164-172); RandomBiasedAI and the scripted AIs are device policies, so the batch hides the state from them
        po_pol = bool(partial_obs) and not selfplay and spec is not None and spec.policy >= M.POLICY_RANDOM_BIASED
        self.b = M.BatchedGameState(utt, maps[0] if same else maps, n, device=device, partial_obs=partial_obs, scripted_ai=scripted or po_pol,
                                    po_policies=po_pol)
        b = self.b
        if selfplay:
            b.set_policy(0, M.POLICY_EXTERNAL)
            b.set_policy(1, M.POLICY_EXTERNAL)
            b.set_issue_order(True)
        else:
            b.set_policy(side, M.POLICY_EXTERNAL)
            b.set_policy(1 - side, spec.policy, spec.pathfinder)
        self.seeds = np.asarray(envs[0::2] if selfplay else envs, dtype=np.int64) + seed0  # the game of environment e is seeded with seed + e
        self.info = _device_array((n, 2, 12), np.int32, emulated)
        b.set_info_output(self.info)
        self.players = [0, 1] if selfplay else [side]
        self.contiguous = bool((np.diff(self.envs) == 1).all()) if len(envs) > 1 else True
        done_mode = getattr(rfs[0], "DONE_MODE", None) if rfs else 3
        # fast path: fused outputs in environment order + auto-reset inside the step launch
        self.fast = (not partial_obs) and done_mode is not None
        self.obs_dtype = np.uint8 if compact else np.int32
        self.fused_masks = bool(fused_masks) and self.fast
        shape = (b.num_planes, b.height, b.width)
        if self.fast:
            self.obs_dev = _device_array((self.E,) + shape, self.obs_dtype, emulated)
            self.obs_host, self._obs_pin = _host_array((self.E,) + shape, self.obs_dtype, emulated)
            self.mask_dev = self.mask_host = self._mask_pin = None
            if self.fused_masks:
                mshape = (self.E, b.height, b.width, (b.mask_width + 7) // 8)
                self.mask_dev = _device_array(mshape, np.uint8, emulated)
                self.mask_host, self._mask_pin = _host_array(mshape, np.uint8, emulated)
            b.set_output_layout(self.obs_dev, self.mask_dev, interleaved=selfplay, side=side)
            b.set_vec_autoreset(done_mode, max_steps)
            self.res_dev = _device_array((n, 4), np.int32, emulated)
            self.res_host, self._res_pin = _host_array((n, 4), np.int32, emulated)
            self.info_host, self._info_pin = _host_array((n, 2, 12), np.int32, emulated)
        else:
            self.obs = {p: _device_array((n,) + shape, np.int32, emulated) for p in self.players}
            b.set_observation_outputs(self.obs.get(0), self.obs.get(1))  # every step() leaves the new observations in self.obs
        self.steps = np.zeros(n, dtype=np.int64)

    # -- fast path -------------------------------------------------------------------------------------------------------
    def stage_and_step(self, action):
        """Actions in (one host -> device copy), one step launch, results / reward facts / observations (/ masks) out -- all queued on the
        batch's stream without waiting."""
        b = self.b
        block = action[self.envs[0]:self.envs[-1] + 1] if self.contiguous else np.ascontiguousarray(action[self.envs])
        if self.selfplay:
            b.vec_step(block, async_copy=True)   # actions in + the one-cycle step: one library call
        else:
            b.set_actions(self.side, block, fill_none_duration=1)
            b.step(1, NO_CAP)
        self._block_keepalive = block
        if self.emulated:
            b.results(self.res_host)
            self.info_host[...] = self.info
            self.obs_host[...] = self.obs_dev
            if self.fused_masks:
                self.mask_host[...] = self.mask_dev
            return
        b.results(self.res_dev)  # a device -> device copy of what the step kernel left
        b.copy_to_host(self.res_host, self.res_dev)  # all queued on the batch's own stream (mrts_batch_copy_to_host); sync() waits
        b.copy_to_host(self.info_host, self.info)
        b.copy_to_host(self.obs_host, self.obs_dev)
        if self.fused_masks:
            b.copy_to_host(self.mask_host, self.mask_dev)

    def download_observations(self):
        """After a reset: observations (and masks) of the current state."""
        b = self.b
        for p in self.players:
            o = b.observe(p, self.obs_dtype)  # host array [n][C][H][W]
            if self.selfplay:
                self.obs_host[p::2] = o
            else:
                self.obs_host[...] = o
            if self.fused_masks:
                mk = b.masks(p, "bits")
                if self.selfplay:
                    self.mask_host[p::2] = mk
                else:
                    self.mask_host[...] = mk

    # -- host-driven path (partially observable batches, custom reward classes) --------------------------------------------
    def observe_all(self):
        for p in self.players:
            self.b.observe(p, np.int32, out=self.obs[p])

    def observations(self):
        self.b.sync()
        return {p: _host(self.obs[p]) for p in self.players}


class JNIGridnetVecClient:
    def __init__(self, a_num_selfplayenvs, a_num_envs, a_max_steps, a_rfs, a_micrortsPath, a_mapPaths, a_ai2s, a_utt, partial_obs=False,
                 device=0, seed=0, compact=False, fused_masks=None):
        from . import _ffi
        assert a_num_selfplayenvs % 2 == 0, "self-play environments come in pairs"
        self.maxSteps, self.utt, self.rfs, self.partialObs, self.mapPaths = a_max_steps, a_utt, list(a_rfs), partial_obs, list(a_mapPaths)
        self.num_selfplay, self.num_envs = a_num_selfplayenvs, a_num_envs
        self.compact = bool(compact)
        self.fused_masks = self.compact if fused_masks is None else bool(fused_masks)
        s1 = a_num_selfplayenvs + a_num_envs
        assert len(self.mapPaths) >= s1 and len(a_ai2s or []) >= a_num_envs
        emulated = "emu" in os.path.basename(getattr(_ffi.lib(), "_name", "") or "")
        cache = {}

        def load(path):
            if isinstance(path, M.PhysicalGameState):
                return path
            full = os.path.join(a_micrortsPath, path) if a_micrortsPath else path
            if full not in cache:
                cache[full] = M.PhysicalGameState.load(full, a_utt)
            return cache[full]

        self._device, self._emulated, self._seed, self._load = device, emulated, seed, load
        self.groups = []
        if a_num_selfplayenvs:
            envs = list(range(a_num_selfplayenvs))
            maps = [load(self.mapPaths[i * 2]) for i in range(a_num_selfplayenvs // 2)]
            self.groups.append(self._make_group(maps, envs, True, None, 0))
        self._bot_specs = list(a_ai2s or [])
        self._bot_groups = {}  # built lazily in reset(): the agent's side comes with `players`
        b0 = self.groups[0].b if self.groups else None
        self.observation = None
        self.reward = np.zeros((s1, len(self.rfs)), dtype=np.float64)
        self.done = np.zeros((s1, len(self.rfs)), dtype=bool)
        self.responses = Responses(None, None, None)
        self._s1 = s1
        self._shape = None if b0 is None else (b0.num_planes, b0.height, b0.width)
        self._packed = None

    def _make_group(self, maps, envs, selfplay, spec, side):
        return _Group(self.utt, maps, envs, selfplay, spec, side, self.partialObs, self._device, self._emulated, self._seed, self.rfs,
                      self.maxSteps, self.compact, self.fused_masks)

    # ---------------------------------------------------------------------------------------------------------------
    def _build_bot_groups(self, players):
        keyed = {}
        for i in range(self.num_envs):
            env = self.num_selfplay + i
            spec = self._bot_specs[i]
            keyed.setdefault((spec.key(), int(players[env])), []).append(env)
        for (skey, side), envs in keyed.items():
            spec = self._bot_specs[envs[0] - self.num_selfplay]
            maps = [self._load(self.mapPaths[e]) for e in envs]
            g = self._make_group(maps, envs, False, spec, side)
            self._bot_groups[(skey, side)] = g
            self.groups.append(g)
        b0 = self.groups[0].b
        self._shape = (b0.num_planes, b0.height, b0.width)

    def _single(self):
        """The one group that covers every environment in order, if there is one: the client's arrays are then that group's."""
        g = self.groups[0] if len(self.groups) == 1 else None
        return g if g is not None and g.fast and g.E == self._s1 and g.contiguous else None

    def _gather_observations(self):
        one = self._single()
        if one is not None:
            self.observation = one.obs_host
            self._packed = one.mask_host
            return
        if self.observation is None:
            self.observation = np.zeros((self._s1,) + self._shape, dtype=np.uint8 if self.compact else np.int32)
        for g in self.groups:
            if g.fast:
                self.observation[g.envs] = g.obs_host
                if g.fused_masks:
                    if self._packed is None:
                        self._packed = np.zeros((self._s1,) + g.mask_host.shape[1:], dtype=np.uint8)
                    self._packed[g.envs] = g.mask_host
                continue
            o = g.observations()
            if g.selfplay:
                self.observation[g.envs[0::2]] = o[0]
                self.observation[g.envs[1::2]] = o[1]
            else:
                self.observation[g.envs] = o[g.side]

    def reset(self, players):
        if self.num_envs and not self._bot_groups:
            self._build_bot_groups(players)
        for g in self.groups:
            g.b.reset(g.seeds)
            g.steps[:] = 0
            if g.fast:
                g.download_observations()
            else:
                g.observe_all()
        self._gather_observations()
        self.reward[:] = 0
        self.done[:] = False
        self.responses.set(self.observation, self.reward, self.done)
        return self.responses

    def gameStep(self, action, players):
        if not isinstance(action, np.ndarray) or action.dtype != np.int32 or not action.flags["C_CONTIGUOUS"]:
            action = np.ascontiguousarray(action, dtype=np.int32)
        assert action.ndim == 3 and action.shape[0] == self._s1 and action.shape[2] == 8, "action = [envs][k][8] vector actions"
        for g in self.groups:
            if g.fast:
                g.stage_and_step(action)
                continue
            b = g.b
            if g.selfplay:
                b.set_actions(0, np.ascontiguousarray(action[g.envs[0::2]]), fill_none_duration=1)
                b.set_actions(1, np.ascontiguousarray(action[g.envs[1::2]]), fill_none_duration=1)
            else:
                b.set_actions(g.side, np.ascontiguousarray(action[g.envs]), fill_none_duration=1)
            b.step(1, NO_CAP)
        for g in self.groups:
            b = g.b
            if g.fast:
                b.sync()
                res, info = g.res_host, g.info_host
            else:
                res, info = b.results(), _host(g.info)
            g.steps += 1
            sides = [(0, g.envs[0::2]), (1, g.envs[1::2])] if g.selfplay else [(g.side, g.envs)]
            for side, envs in sides:
                for j, rf in enumerate(self.rfs):
                    r, d = rf.compute(info[:, side], res, side)
                    self.reward[envs, j] = r
                    self.done[envs, j] = d
            # auto-reset (JNIGridnetVecClient.java:244-262,272-286): terminal reward/done are kept, done[0] is forced
            if g.fast:
                finished = (res[:, 2] & 2) != 0  # the step kernel restarted these games before it wrote their observations
                if finished.any():
                    for side, envs in sides:
                        self.done[np.asarray(envs)[finished], 0] = True
                continue
            first = g.envs[0::2] if g.selfplay else g.envs
            finished = self.done[first, 0] | (g.steps >= self.maxSteps)
            if finished.any():
                b.restart_masked(finished.astype(np.uint8))
                g.steps[finished] = 0
                for side, envs in sides:
                    self.done[np.asarray(envs)[finished], 0] = True
                g.observe_all()
        self._gather_observations()
        self.responses.set(self.observation, self.reward, self.done)
        return self.responses

    def getMasks(self, player):
        one = self._single()
        if one is not None and one.selfplay and not self._emulated:
            # dense int32 masks written by the device straight in environment order into a reused device array, one copy into a reused
            # pinned array (the reference's getMasks also returns buffers it owns)
            g, b = one, one.b
            if getattr(g, "dense_dev", None) is None:
                shape = (g.E, b.height, b.width, b.mask_width)
                g.dense_dev = _device_array(shape, np.int32, False)
                g.dense_host, g._dense_pin = _host_array(shape, np.int32, False)
            per_env = b.height * b.width * b.mask_width
            flat = g.dense_dev.view(-1)
            b.masks(0, np.int32, out=flat[0:])           # player 0 -> environments 0, 2, 4, ... (output stride 2)
            b.masks(1, np.int32, out=flat[per_env:])     # player 1 -> environments 1, 3, 5, ...
            b.copy_to_host(g.dense_host, g.dense_dev)
            b.sync()
            return g.dense_host
        out = None
        for g in self.groups:
            if g.selfplay:
                m0, m1 = g.b.masks(0), g.b.masks(1)
                if out is None:
                    out = np.zeros((self._s1,) + m0.shape[1:], dtype=np.int32)
                out[g.envs[0::2]] = m0
                out[g.envs[1::2]] = m1
            else:
                m = g.b.masks(g.side)
                if out is None:
                    out = np.zeros((self._s1,) + m.shape[1:], dtype=np.int32)
                out[g.envs] = m
        return out

    def getMasksPacked(self):
        """The masks of getMasks bit-packed, [envs][H][W][(mask width + 7) // 8] uint8 (element j of a cell in bit j & 7 of byte j >> 3),
        as the last reset / gameStep left them -- needs fused_masks (implied by compact=True): they arrive with the observations."""
        assert self.fused_masks and self._packed is not None, "construct the client with fused_masks=True (or compact=True)"
        return self._packed

    def close(self):
        for g in self.groups:
            g.b.close()
        self.groups = []
