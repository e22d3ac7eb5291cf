#!/bin/bash
# config 3 again after the A* changes: launch list + one --set full capture of the lean rush kernel (label r2z)
mkdir -p gpurun_out
rm -f gpurun_out/traffic_r2z.json
tools/profile2.sh r2z cfg3 k_step_fixed_24x24_rush 14 65536 --workload scripted --steps 2 --warmup 3
python - <<'PY'
import json
t=json.load(open('gpurun_out/traffic_r2z.json'))
for k,v in t.items(): print(k, {kk:(round(vv,3) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk!='source'})
PY
