#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/s6_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s6_pytest.log
tail -3 gpurun_out/s6_pytest.log
B="python bench.py --no-cpu-baseline --no-e2e --no-secondary --steps 40 --warmup 5"
$B --workload obs > gpurun_out/s6_obs.json 2>gpurun_out/s6_obs.err
$B --workload obs --with-masks > gpurun_out/s6_obsm.json 2>gpurun_out/s6_obsm.err
$B > gpurun_out/s6_self.json 2>gpurun_out/s6_self.err
$B --workload vec > gpurun_out/s6_vec.json 2>gpurun_out/s6_vec.err
$B --workload scripted --steps 10 > gpurun_out/s6_scr.json 2>gpurun_out/s6_scr.err
for f in obs obsm self vec scr; do python - $f <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/s6_%s.json'%f).read().strip().splitlines()[-1])
    r=d['roofline']
    print(f, '%.4g'%d['value'], 'ms/step %.3f'%d['ms_per_step'], 'kernel_ms %.3f'%r['mean_launch_ms'], r['kernel'], 'frac %.3f'%r['frac'], 'achieved %.0f GB/s'%r['achieved'], 'dram_frac %.3f'%r['dram_frac'], d.get('e2e') and d['e2e'].get('value'))
except Exception as e:
    print(f, 'FAILED', e); print(open('gpurun_out/s6_%s.err'%f).read()[-800:])
PY
done
python - <<'PY'
# MCTS throughput: 1024 searches of 100 iterations on 16x16
import time, numpy as np, sys
sys.path.insert(0, '.')
import microrts_b200 as M
from microrts_b200 import search as S
utt = M.UnitTypeTable(1, 1)
for T in (256, 2048):
    b = M.BatchedGameState(utt, M.maps.standard_map("16x16/basesWorkers16x16", utt), T)
    b.set_policy(0, M.POLICY_RANDOM_BIASED); b.set_policy(1, M.POLICY_RANDOM_BIASED)
    b.reset(np.arange(T, dtype=np.int64)); b.step(200, 3000); b.cycle_to_decision()
    s = S.NaiveMCTS(b, 0, max_nodes_per_tree=102)
    s.iterate(2); b.sync()
    t0 = time.perf_counter(); s.iterate(98); dt = time.perf_counter() - t0
    print("NaiveMCTS %d searches x 98 iterations: %.3f s = %.3g playouts/s (%.2f ms per lockstep iteration)" % (T, dt, T * 98 / dt, dt / 98 * 1e3))
    s.close(); b.close()
PY
