#!/bin/bash
mkdir -p gpurun_out
tools/profile2.sh r2a cfg5 k_step_fast_obs 3 65536 --workload obs --steps 2 --warmup 3
tools/profile2.sh r2a cfg5_masks k_step_fast_obs 3 65536 --workload obs --with-masks --steps 2 --warmup 3
tools/profile2.sh r2a cfg3 k_step_fixed_24x24_rush 14 65536 --workload scripted --steps 2 --warmup 3
tools/profile2.sh r2a cfg2 "k_fixed" 33 65536 --steps 2 --warmup 3
cat gpurun_out/traffic_r2a.json
