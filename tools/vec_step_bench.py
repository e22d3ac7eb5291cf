#!/usr/bin/env python3
"""Micro-benchmark of the RL-style step (the JNIGridnetVecClient flow on the device): per step H2D of vector actions for both
players, one k_step launch (decode, issueSafe x2 in self-play order, cycle, reward facts, both observations), D2H of results
and reward facts.  usage: python tools/vec_step_bench.py [games] [map key]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import golden_io, microrts_b200 as M, parity as P
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
key = sys.argv[2] if len(sys.argv) > 2 else "16x16/basesWorkers16x16"
maps = golden_io.load_maps()
utt = M.UnitTypeTable(1, 1)
b = M.BatchedGameState(utt, M.PhysicalGameState.fromXML(P.map_to_xml(maps[key]), utt), n)
b.set_policy(0, M.POLICY_EXTERNAL); b.set_policy(1, M.POLICY_EXTERNAL); b.set_issue_order(True)
W, H = b.width, b.height
obs = [torch.empty((n, 6, H, W), dtype=torch.int32, device="cuda") for _ in range(2)]
info = torch.zeros((n, 2, 12), dtype=torch.int32, device="cuda")
b.set_observation_outputs(obs[0], obs[1]); b.set_info_output(info)
K = 16
rng = np.random.default_rng(0)
acts = [np.zeros((n, K, 8), dtype=np.int32) for _ in range(2)]
for a in acts:
    a[:, :, 0] = rng.integers(0, W * H, size=(n, K)); a[:, :, 1] = rng.integers(0, 6, size=(n, K))
    a[:, :, 2:6] = rng.integers(0, 4, size=(n, K, 4)); a[:, :, 6] = rng.integers(1, 7, size=(n, K)); a[:, :, 7] = rng.integers(0, 49, size=(n, K))
def step():
    b.set_actions(0, acts[0], fill_none_duration=1); b.set_actions(1, acts[1], fill_none_duration=1)
    b.step(1, 1 << 30)
    r = b.results(); i = info.cpu()
    return r
for _ in range(20): step()
b.sync(); t0 = time.perf_counter()
reps = 100
for _ in range(reps): step()
b.sync(); dt = (time.perf_counter() - t0) / reps
print("%s x %d games: %.3f ms per RL step (%.2e env-steps/s incl. host copies of actions, results, reward facts)" % (key, n, dt * 1e3, n / dt))
st = b.stats(); print(st)
