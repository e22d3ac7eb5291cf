#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s4_pytest.log
tail -3 gpurun_out/s4_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/s4_bench.json 2> gpurun_out/s4_bench.err; echo "bench rc=$?"
tail -5 gpurun_out/s4_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s4_bench.json').read().strip().splitlines()[-1])
print('headline', d['value'], d['ms_per_step'], 'e2e', d['e2e'] and d['e2e']['value'], 'roofline', {k:d['roofline'].get(k) for k in ('kernel','frac','dram_frac','issue_frac')})
for k,v in (d.get('secondary') or {}).items():
    print(k, v.get('value'), v.get('unit'), v.get('ms_per_step'), v.get('error'), (v.get('roofline') or {}).get('frac'), (v.get('roofline') or {}).get('kernel'), (v.get('cpu_baseline') or {}).get('value'), {kk:vv for kk,vv in (v.get('config') or {}).items() if kk in ('mean_rollout_cycles','contact_fraction','rollouts_per_s')})
    if k=='vec': print('   e2e', v.get('e2e'), '\n   ref', v.get('e2e_reference_layout'))
PY
tools/profile2.sh r2a cfg5 k_step_fast_obs 34 65536 --workload obs --steps 2 --warmup 3
tools/profile2.sh r2a cfg5_masks k_step_fast_obs 34 65536 --workload obs --with-masks --steps 2 --warmup 3
tools/profile2.sh r2a vec "k_step" 6 16384 --workload vec --steps 2 --warmup 3
ls -la gpurun_out | grep r2a | head -30
