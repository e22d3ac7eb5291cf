#!/bin/bash
mkdir -p gpurun_out
tools/profile2.sh r2c cfg3 k_step_fixed_24x24_rush 14 65536 --workload scripted --steps 2 --warmup 3
ls -la gpurun_out | grep r2c
