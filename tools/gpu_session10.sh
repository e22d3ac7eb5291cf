#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_benchmark_shape.py tests/test_mcts.py tests/test_vec_client.py -m gpu -x -q 2>&1 | tail -4
python - <<'PY'
import time, numpy as np, sys
sys.path.insert(0, '.')
import microrts_b200 as M
from microrts_b200 import search as S
utt = M.UnitTypeTable(1, 1)
for T in (256, 2048, 8192):
    b = M.BatchedGameState(utt, M.maps.standard_map("16x16/basesWorkers16x16", utt), T)
    b.set_policy(0, M.POLICY_RANDOM_BIASED); b.set_policy(1, M.POLICY_RANDOM_BIASED)
    b.reset(np.arange(T, dtype=np.int64)); b.step(200, 3000); b.cycle_to_decision()
    s = S.NaiveMCTS(b, 0, max_nodes_per_tree=102)
    s.iterate(2); b.sync()
    t0 = time.perf_counter(); s.iterate(98); dt = time.perf_counter() - t0
    print("NaiveMCTS %d searches x 98 iterations: %.3f s = %.3g playouts/s (%.2f ms per lockstep iteration)" % (T, dt, T * 98 / dt, dt / 98 * 1e3))
    s.close(); b.close()
PY
python bench.py --no-cpu-baseline --no-e2e --no-secondary --steps 40 --warmup 5 --workload vec 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('vec', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e_reference_layout']['value'])"
