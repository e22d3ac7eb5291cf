#!/bin/bash
# quick A/B of the short-step workloads
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-e2e --no-secondary --steps 40 --warmup 5"
for w in "obs" "obs --with-masks" "vec"; do
  n=$(echo $w | tr -d ' -')
  $B --workload $w > gpurun_out/q_$n.json 2>gpurun_out/q_$n.err
  python - $n <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/q_%s.json'%f).read().strip().splitlines()[-1])
    r=d['roofline']
    print(f, '%.4g'%d['value'], 'ms/step %.3f'%d['ms_per_step'], 'kernel_ms %.3f'%r['mean_launch_ms'], r['kernel'], 'frac %.3f'%r['frac'], 'achieved %.0f GB/s'%r['achieved'], d.get('e2e') and d['e2e'].get('value'))
except Exception as e:
    print(f, 'FAILED', e); print(open('gpurun_out/q_%s.err'%f).read()[-800:])
PY
done
