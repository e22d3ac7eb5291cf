#!/bin/bash
L=${1:-x}
CMD="python bench.py --workload scripted --steps 2 --warmup 6 --no-cpu-baseline --prewarm-seconds 0"
$CMD > gpurun_out/plain_scr_$L.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_scr_$L.log; exit 1; }
ncu --set full --clock-control none --import-source on -k "regex:k_step" -s 7 -c 1 -f -o gpurun_out/prof_scr_$L $CMD > gpurun_out/ncu_full_scr_$L.log 2>&1
ncu -i gpurun_out/prof_scr_$L.ncu-rep --page raw --csv > gpurun_out/prof_scr_${L}_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_scr_$L.ncu-rep --page source --print-source sass --csv > gpurun_out/prof_scr_${L}_sass.csv 2>/dev/null
ls -la gpurun_out/*scr_$L*
