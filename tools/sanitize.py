#!/usr/bin/env python3
"""Small pass over every kernel of the library, sized for a memory checker.

On a GPU box with compute-sanitizer available:
    compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize.py
(compute-sanitizer is closed on the pool this repository is developed on.)  Without one, the same kernel sources run on the warp
emulator (tests/emu) built with AddressSanitizer + UndefinedBehaviorSanitizer -- every "device" buffer is a heap block there:
    LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0 \
        MRTS_EMU=1 MRTS_EMU_ASAN=1 python tools/sanitize.py          # or: ... python -m pytest tests -m gpu

Each section launches one kernel kind (fast self-play, fixed-layout copies, the generic scripted kernel with every policy family and
pathfinder, the lean rush kernel, partially observable policies, rollouts, observation / mask / evaluation / pathfinding operators,
the fused step + observation kernel with masks, the JNIGridnetVecClient step with in-kernel reset, game copies, the host searches).
Prints the kernels launched; the checker's own report follows."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
EMU = os.environ.get("MRTS_EMU", "0") == "1"  # dry run of this script on the warp emulator (tests/emu)
if EMU:
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "emu"))
    import emu_backend
    emu_backend.use_emulator()
else:
    import torch  # device buffers for the fused outputs

import microrts_b200 as M
from microrts_b200 import search as S
from microrts_b200 import vec_client as V

utt = M.UnitTypeTable(1, 1)
kernels = set()


def note(b):
    kernels.add(b.last_kernel)


def dev_zeros(shape, dtype):
    if EMU:
        return np.zeros(shape, dtype=dtype)
    return torch.zeros(shape, dtype=getattr(torch, np.dtype(dtype).name), device="cuda")


def batch(key, n, **kw):
    b = M.BatchedGameState(utt, M.maps.standard_map(key, utt), n, **kw)
    b.reset(np.arange(n, dtype=np.int64) + 7)
    return b


# 1. RandomBiasedAI self-play: fixed-layout copies (8x8, 16x16) and the generic fast kernel (12x12), with auto-reset
for key in ("8x8/basesWorkers8x8", "16x16/basesWorkers16x16", "12x12/basesWorkers12x12"):
    b = batch(key, 96)
    b.set_policy(0, M.POLICY_RANDOM_BIASED); b.set_policy(1, M.POLICY_RANDOM_BIASED)
    b.set_auto_reset(True)
    for _ in range(3):
        b.step(150, 400); note(b)
    b.stats(); b.close()

# 2. scripted policies: lean rush kernels, the generic kernel with every family and pathfinder, partially observable policies
SCRIPTED = [("16x16/basesWorkers16x16", "WORKER_RUSH", "LIGHT_RUSH", M.PF_ASTAR, False),
            ("24x24/basesWorkers24x24", "LIGHT_RUSH", "WORKER_RUSH", M.PF_ASTAR, False),
            ("16x16/basesWorkers16x16", "CRUSH_V2", "EMR_DETERMINISTICO", M.PF_ASTAR, False),
            ("8x8/basesWorkers8x8", "CRUSH_V1", "WORKER_RUSH_PP", M.PF_BFS, False),
            ("16x16/TwoBasesBarracks16x16", "LIGHT_DEFENSE", "PO_HEAVY_RUSH", M.PF_BFS, True),
            ("16x16/basesWorkers16x16", "RANGED_RUSH", "HEAVY_DEFENSE", M.PF_FLOODFILL, False),
            ("BWDistantResources32x32", "CRUSH_V2", "RANDOM_BIASED", M.PF_GREEDY, True)]
for key, p0, p1, pf, po in SCRIPTED:
    b = batch(key, 24, scripted_ai=True, po_policies=po)
    b.set_policy(0, getattr(M, "POLICY_" + p0), pf); b.set_policy(1, getattr(M, "POLICY_" + p1), pf)
    for _ in range(4):
        b.step(300, 3000); note(b)
    b.export(); b.close()

# 3. operators on a mid-game state: rollouts (fully and partially observable), evaluation, observation planes, masks (dense and
#    bit-packed), ordered unit-action lists, the MCTS node loop, pathfinding queries, export / import, copies, masked reset
b = batch("16x16/basesWorkers16x16", 64, scripted_ai=True)
b.set_policy(0, M.POLICY_RANDOM_BIASED); b.set_policy(1, M.POLICY_RANDOM_BIASED)
b.step(300, 3000); note(b)
for observer in (-1, 0):
    b.rollout(depth=60, rollouts_per_game=3, observer=observer); note(b)
b.evaluate(0, 0); b.evaluate(1, 1, observer=1); note(b)
for pl in (0, 1):
    b.observe(pl); note(b)
    b.observe(pl, dtype=np.uint8)
    b.masks(pl); note(b)
    b.masks(pl, dtype="bits")
b.cycle_to_decision(); note(b)
b.unit_actions(0); note(b)
cells = np.zeros(64, dtype=np.int32); targets = np.full(64, 255, dtype=np.int32); ranges = np.ones(64, dtype=np.int32)
for pf in (M.PF_ASTAR, M.PF_BFS, M.PF_GREEDY):
    b.find_path(pf, cells, targets, ranges); note(b)
st = b.export()
b2 = batch("16x16/basesWorkers16x16", 64, scripted_ai=True)
b2.import_(st)
b2.copy_games(b, np.arange(64, dtype=np.int64)[::-1].copy(), None); note(b2)
mask = np.zeros(64, dtype=np.uint8); mask[::3] = 1
b2.reset_masked(mask, np.arange(64, dtype=np.int64))
b2.cycle(5); note(b2)
b2.close()

# 4. host searches over the batch (getPlayerActions, NaiveMCTS, UCT)
b.close()
b = batch("16x16/basesWorkers16x16", 16)
b.set_policy(0, M.POLICY_RANDOM_BIASED); b.set_policy(1, M.POLICY_RANDOM_BIASED)
b.step(200, 3000); b.cycle_to_decision()
srch = S.NaiveMCTS(b, 0, max_nodes_per_tree=12)
srch.iterate(10); srch.close()
srch = S.UCT(b, 1, max_nodes_per_tree=12)
srch.iterate(10); srch.close()
b.close()

# 5. the fused step + observation kernel (slim layout, bulk stores) with both players' planes and bit-packed masks, 64x64
for key, n in (("GardenOfWar64x64", 6), ("16x16/basesWorkers16x16", 40)):
    b = batch(key, n)
    b.set_policy(0, M.POLICY_RANDOM_BIASED); b.set_policy(1, M.POLICY_RANDOM_BIASED)
    b.step(60, 3000)
    for dt in (np.uint8, np.int32):
        o = [dev_zeros((n, b.num_planes, b.height, b.width), dt) for _ in range(2)]
        m = [dev_zeros((n, b.height, b.width, (b.mask_width + 7) // 8), np.uint8) for _ in range(2)]
        b.set_observation_outputs(o[0], o[1]); b.set_mask_outputs(m[0], m[1])
        for _ in range(6):
            b.step(1, 3000); note(b)
        b.sync()
        b.set_mask_outputs(None, None)
        for _ in range(3):
            b.step(1, 3000); note(b)
        b.sync()
    b.set_observation_outputs(None, None)
    b.close()

# 6. the JNIGridnetVecClient flow: self-play pairs and bot environments, in-kernel reset, compact and reference layouts
from microrts_b200 import rewards as R
pgs8 = M.maps.standard_map("8x8/basesWorkers8x8", utt)
for compact in (True, False):
    vc = V.JNIGridnetVecClient(4, 2, 60, [R.WinLossRewardFunction(), R.ResourceGatherRewardFunction(), R.AttackRewardFunction()], "", [pgs8] * 6,
                               [V.ai.WorkerRush(utt), V.ai.CRush_V1(utt)], utt, partial_obs=False, seed=5, compact=compact)
    vc.reset([0] * 6)
    acts = np.zeros((6, 8, 8), dtype=np.int32)
    for _ in range(70):
        vc.gameStep(acts, [0] * 6)
        if compact:
            vc.getMasksPacked()
        else:
            vc.getMasks(0)
    vc.close()

print("kernels launched:", sorted(kernels))
print("sanitize.py: done")
