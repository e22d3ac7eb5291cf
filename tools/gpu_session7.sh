#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/s7_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s7_pytest.log
tail -3 gpurun_out/s7_pytest.log
B="python bench.py --no-cpu-baseline --no-e2e --no-secondary --warmup 5"
$B --workload scripted --steps 10 > gpurun_out/s7_scr.json 2>gpurun_out/s7_scr.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s7_scr.json').read().strip().splitlines()[-1]); r=d['roofline']
print('scr', '%.4g'%d['value'], 'kernel_ms %.3f'%r['mean_launch_ms'], r['kernel'])
PY
