#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_engine_parity.py tests/test_benchmark_shape.py -m gpu -x -q -k "observation_kernel or fused or cfg5 or selfplay or wide_differential_full" 2>&1 | tail -3
bash tools/gpu_quick.sh
B="python bench.py --no-cpu-baseline --no-e2e --no-secondary --steps 40 --warmup 5"
$B > gpurun_out/s8_self.json 2>gpurun_out/s8_self.err; python -c "
import json; d=json.loads(open('gpurun_out/s8_self.json').read().strip().splitlines()[-1]); print('self %.4g'%d['value'], d['roofline']['mean_launch_ms'])"
