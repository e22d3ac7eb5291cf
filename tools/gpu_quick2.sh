#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vec_client.py tests/test_engine_parity.py -m gpu -x -q -k "vec or masks or scripted_policies" 2>&1 | tail -3
python bench.py --no-cpu-baseline --no-e2e --no-secondary --steps 40 --warmup 5 --workload vec > gpurun_out/q2_vec.json 2>gpurun_out/q2_vec.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/q2_vec.json').read().strip().splitlines()[-1])
print('vec', d['value'], d['ms_per_step']); print(' e2e', d['e2e']); print(' ref', d['e2e_reference_layout'])
PY
tail -3 gpurun_out/q2_vec.err
