#!/usr/bin/env python3
"""Where the time of bench.py's e2e loop goes: per step, host time spent submitting (reset_masked + step), waiting in results()
and in the numpy bookkeeping, with the batch split into 2 or more parts.  usage: python tools/e2e_diag.py [parts]"""
import os, sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
import golden_io, microrts_b200 as M, parity as P
maps = golden_io.load_maps(); utt = M.UnitTypeTable(1, 1)
pgs = M.PhysicalGameState.fromXML(P.map_to_xml(maps["16x16/basesWorkers16x16"]), utt)
n, C, MAXC = 65536, 100, 3000
nh = int(sys.argv[1]) if len(sys.argv) > 1 else 2
seeds = np.arange(n, dtype=np.int64)
halves = []
cuts = [n * i // nh for i in range(nh + 1)]
for lo, hi in zip(cuts[:-1], cuts[1:]):
    hb = M.BatchedGameState(utt, pgs, hi - lo)
    hb.set_policy(0, M.POLICY_RANDOM_BIASED); hb.set_policy(1, M.POLICY_RANDOM_BIASED); hb.set_auto_reset(False); hb.reset(seeds[lo:hi])
    h = dict(b=hb, mask=torch.zeros(hi - lo, dtype=torch.uint8).pin_memory(), seeds=torch.from_numpy(seeds[lo:hi].copy()).pin_memory(), res=torch.zeros((hi - lo, 4), dtype=torch.int32).pin_memory(), ep=0)
    h["mask_np"], h["seeds_np"], h["res_np"] = h["mask"].numpy(), h["seeds"].numpy(), h["res"].numpy()
    halves.append(h)
T = dict(submit=0.0, wait=0.0, host=0.0)
def submit(h):
    t = time.perf_counter()
    h["b"].reset_masked(h["mask_np"], h["seeds_np"]); h["b"].step(C, MAXC)
    T["submit"] += time.perf_counter() - t
def collect(h):
    t = time.perf_counter()
    h["b"].results(h["res_np"])
    t2 = time.perf_counter(); T["wait"] += t2 - t
    r = h["res_np"]; done = (r[:, 2] != 0) | (r[:, 0] >= MAXC); h["mask_np"][:] = done
    if done.any(): h["ep"] += 1; h["seeds_np"][done] += n * h["ep"]
    s = int(r[:, 0].sum())
    T["host"] += time.perf_counter() - t2
for h in halves: submit(h)
for _ in range(30):
    for h in halves: collect(h); submit(h)
for k in T: T[k] = 0.0
torch.cuda.synchronize(); t0 = time.perf_counter(); K = 150
for _ in range(K):
    for h in halves: collect(h); submit(h)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(os.environ.get("MRTS_CUDA_LIB"), "parts", nh, "ms/step %.3f" % (dt / K * 1e3), {k: "%.3f" % (v / K * 1e3) for k, v in T.items()})
