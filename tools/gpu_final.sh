#!/bin/bash
# the committed library: whole gpu suite, smoke, the default bench line and the reference arm
mkdir -p gpurun_out
timeout 3000 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python __graft_entry__.py --smoke 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 > gpurun_out/final_n1.json 2> gpurun_out/final_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/final_n1.json').read().strip().splitlines()[-1])
print('headline %.4g e2e %.4g frac %.3f cpu %.4g launches %s clocks %s'%(d['value'],d['e2e']['value'],d['roofline']['frac'],d['cpu_baseline']['value'],d['gpu_launches'],d['clocks']))
for k,v in (d.get('secondary') or {}).items(): print(' ',k,'%.4g'%v.get('value'),v.get('error'))
r=json.loads(open('gpurun_out/final_ref.json').read().strip().splitlines()[-1]); print('reference arm %.4g %s'%(r['value'], r['cpu_baseline']))
PY
