#!/bin/bash
# round-2 final captures (label r2z): one launch list + one --set full capture per configuration, summaries for profiles/
mkdir -p gpurun_out
rm -f gpurun_out/traffic_r2z.json
tools/profile2.sh r2z cfg2 "k_fixed" 33 65536 --steps 2 --warmup 3
tools/profile2.sh r2z cfg1 "k_fixed" 33 65536 --workload cfg1 --steps 2 --warmup 3
tools/profile2.sh r2z cfg3 k_step_fixed_24x24_rush 14 65536 --workload scripted --steps 2 --warmup 3
tools/profile2.sh r2z cfg4 "k_fixed" 63 16384 --workload rollout --steps 2 --warmup 3
tools/profile2.sh r2z cfg5 k_step_fast_obs 3 65536 --workload obs --steps 2 --warmup 3
tools/profile2.sh r2z cfg5_masks k_step_fast_obs 3 65536 --workload obs --with-masks --steps 2 --warmup 3
tools/profile2.sh r2z vec "k_step" 6 16384 --workload vec --steps 2 --warmup 3
tools/write_ceiling > gpurun_out/r2z_write_ceiling.json
python - <<'PY'
import json
t=json.load(open('gpurun_out/traffic_r2z.json'))
for k,v in t.items(): print(k, {kk:(round(vv,3) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk!='source'})
PY
