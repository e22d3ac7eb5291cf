#!/bin/bash
# round-2 session 1: parity suite on the device, write-only bandwidth ceiling, baseline numbers of the secondary workloads
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/s1_gpu.txt
tools/write_ceiling > gpurun_out/s1_write_ceiling.json 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s1_pytest.log
tail -3 gpurun_out/s1_pytest.log
python tools/vec_step_bench.py 16384 > gpurun_out/s1_vec.log 2>&1
python tools/vec_step_bench.py 65536 >> gpurun_out/s1_vec.log 2>&1
cat gpurun_out/s1_write_ceiling.json; cat gpurun_out/s1_vec.log
