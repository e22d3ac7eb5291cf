#!/usr/bin/env python3
"""Micro-benchmark: HBM write bandwidth of the observation emitters on one GPU (k_observe alone, and the fused step kernel
with 0 cycles), GardenOfWar64x64, uint8 planes.  usage: python tools/obs_bw.py [games]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import golden_io, microrts_b200 as M, parity as P
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
maps = golden_io.load_maps()
utt = M.UnitTypeTable(1, 1)
b = M.BatchedGameState(utt, M.PhysicalGameState.fromXML(P.map_to_xml(maps["GardenOfWar64x64"]), utt), n)
b.set_policy(0, M.POLICY_RANDOM_BIASED); b.set_policy(1, M.POLICY_RANDOM_BIASED)
b.step(200, 3000)
o = [torch.empty((n, 6, 64, 64), dtype=torch.uint8, device="cuda") for _ in range(2)]
from microrts_b200 import _ffi
stream = torch.cuda.ExternalStream(_ffi.lib().mrts_batch_stream(b._h))
def timeit(fn, reps=20):
    for _ in range(3): fn()
    b.sync()
    a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(reps): fn()
    z.record(stream); b.sync()
    return a.elapsed_time(z) / reps
ms = timeit(lambda: b.observe(0, np.uint8, out=o[0]))
print("k_observe: %.3f ms, %.0f GB/s written" % (ms, n * 6 * 4096 / ms / 1e6))
b.set_observation_outputs(o[0], o[1])
ms = timeit(lambda: b.step(0, 3000))
print("fused, 0 cycles (load + rebuild + store + 2 observations): %.3f ms, %.0f GB/s written" % (ms, 2 * n * 6 * 4096 / ms / 1e6))
ms = timeit(lambda: b.step(1, 1 << 30))
print("fused, 1 cycle: %.3f ms, %.0f GB/s written" % (ms, 2 * n * 6 * 4096 / ms / 1e6))
x = torch.empty(n * 6 * 4096 * 2, dtype=torch.uint8, device="cuda")
s2 = torch.cuda.current_stream()
for _ in range(3): x.zero_()
torch.cuda.synchronize()
a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20): x.zero_()
z.record(); torch.cuda.synchronize()
ms = a.elapsed_time(z) / 20
print("torch memset of the same bytes: %.3f ms, %.0f GB/s" % (ms, x.numel() / ms / 1e6))
