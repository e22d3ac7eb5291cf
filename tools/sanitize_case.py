#!/usr/bin/env python3
"""A small run through every kernel, for compute-sanitizer (memcheck / racecheck): a few games of each configuration."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import golden_io, microrts_b200 as M, parity as P
maps = golden_io.load_maps()
utt = M.UnitTypeTable(1, 1)
def pgs(k): return M.PhysicalGameState.fromXML(P.map_to_xml(maps[k]), utt)
n = 24
b = M.BatchedGameState(utt, pgs("16x16/basesWorkers16x16"), n)
b.set_policy(0, M.POLICY_RANDOM_BIASED); b.set_policy(1, M.POLICY_RANDOM_BIASED)
o = [torch.zeros((n, 6, 16, 16), dtype=torch.uint8, device="cuda") for _ in range(2)]
b.step(400, 3000)                       # k_step_fast
b.set_observation_outputs(o[0], o[1]); b.step(50, 3000); b.set_observation_outputs(None, None)   # k_step_fast_obs
b.observe(0); b.masks(1); b.masks(0, "bits")       # k_observe, k_step (masks)
b.rollout(depth=60, rollouts_per_game=3, observer=0); b.rollout(depth=60, rollouts_per_game=2)   # k_rollout
b.set_auto_reset(True); b.step(3000, 3000); b.step(100, 3000)
b.set_policy(1, M.POLICY_EXTERNAL); info = torch.zeros((n, 2, 12), dtype=torch.int32, device="cuda"); b.set_info_output(info)
acts = np.zeros((n, 4, 8), dtype=np.int32); acts[:, :, 0] = np.random.default_rng(0).integers(0, 256, (n, 4)); acts[:, :, 1] = 1
b.set_actions(1, acts); b.step(1, 3000); b.results(); b.export(); b.close()    # k_step generic, external
for key in ("8x8/basesWorkers8x8", "GardenOfWar64x64"):
    s = M.BatchedGameState(utt, pgs(key), 6, scripted_ai=True)
    s.set_policy(0, M.POLICY_LIGHT_RUSH); s.set_policy(1, M.POLICY_WORKER_RUSH, M.PF_BFS)
    s.step(300, 3000); s.sync(); s.close()               # scripted: A* scratch in shared / global memory
po = M.BatchedGameState(utt, pgs("8x8/basesWorkers8x8"), 8, partial_obs=True)
po.set_policy(0, M.POLICY_RANDOM_BIASED); po.set_policy(1, M.POLICY_RANDOM_BIASED); po.step(200, 3000); po.observe(1); po.sync(); po.close()
# partially observable games (po_hide / po_unhide, exploration), the defenses, GreedyPathFinding
for key, p0, p1, pf in (("16x16/basesWorkers16x16", M.POLICY_PO_LIGHT_RUSH, M.POLICY_PO_WORKER_RUSH, M.PF_ASTAR),
                        ("8x8/basesWorkers8x8", M.POLICY_RANDOM_BIASED, M.POLICY_PO_RANGED_RUSH, M.PF_GREEDY),
                        ("GardenOfWar64x64", M.POLICY_PO_HEAVY_RUSH, M.POLICY_LIGHT_DEFENSE, M.PF_BFS)):
    s = M.BatchedGameState(utt, pgs(key), 6, scripted_ai=True, po_policies=True)
    s.set_policy(0, p0, pf); s.set_policy(1, p1, pf)
    s.step(400, 3000); s.sync(); s.close()
s = M.BatchedGameState(utt, pgs("16x16/basesWorkers16x16"), 6, scripted_ai=True)
s.set_policy(0, M.POLICY_WORKER_DEFENSE, M.PF_GREEDY); s.set_policy(1, M.POLICY_WORKER_RUSH_PP); s.step(600, 3000); s.sync(); s.close()
print("sanitize case done")
