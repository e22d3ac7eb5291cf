#!/bin/bash
# usage: tools/profile_other.sh LABEL -- ncu --set full of the observation-emitting kernel (config 5)
L=${1:-x}
for W in "obs:k_step_fast_obs:--workload obs --steps 2 --warmup 6"; do
  name=${W%%:*}; rest=${W#*:}; kern=${rest%%:*}; flags=${rest#*:}
  CMD="python bench.py $flags --no-cpu-baseline --prewarm-seconds 0"
  $CMD > gpurun_out/plain_${name}_$L.log 2>&1 || { echo "plain $name run failed"; continue; }
  ncu --set full --clock-control none -k "regex:$kern" -s 4 -c 1 -f -o gpurun_out/prof_${name}_$L $CMD > gpurun_out/ncu_full_${name}_$L.log 2>&1
  ncu -i gpurun_out/prof_${name}_$L.ncu-rep --page raw --csv > gpurun_out/prof_${name}_${L}_raw.csv 2>/dev/null
done
ls -la gpurun_out/*_$L*raw.csv
