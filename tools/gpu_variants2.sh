#!/bin/bash
timeout 900 python -m pytest tests/test_engine_parity.py tests/test_benchmark_shape.py -m gpu -x -q -k "observation_kernel or fused or cfg5" 2>&1 | tail -2
B="python bench.py --no-cpu-baseline --no-e2e --no-secondary --steps 40 --warmup 5 --workload obs"
for c in 0 1; do
  for m in "" "--with-masks"; do
    MRTS_DBG_OBS_CTAS=$c $B $m 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ctas=[$c] masks=[$m] kernel_ms %.3f' % d['roofline']['mean_launch_ms'], d['roofline']['achieved'])"
  done
done
