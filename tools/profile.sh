#!/bin/bash
# usage: tools/profile.sh LABEL   (on the GPU box; writes gpurun_out/*_LABEL.*)
# 1. the benchmark command without ncu (must exit 0)  2. the launch list  3. one `--set full` capture of k_step_fast
L=${1:-x}
CMD="python bench.py --steps 2 --warmup 15 --no-cpu-baseline --no-e2e --prewarm-seconds 0"
$CMD > gpurun_out/plain_$L.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$L.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$L.csv $CMD > gpurun_out/ncu_$L.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:k_fixed|k_step_fast" -s 16 -c 1 -f -o gpurun_out/prof_$L $CMD > gpurun_out/ncu_full_$L.log 2>&1
ncu -i gpurun_out/prof_$L.ncu-rep --page raw --csv > gpurun_out/prof_${L}_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_$L.ncu-rep --page source --print-source sass --csv > gpurun_out/prof_${L}_sass.csv 2>/dev/null
ls -la gpurun_out/*_$L*
