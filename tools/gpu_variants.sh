#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-e2e --no-secondary --steps 40 --warmup 5 --workload obs"
for v in "" nogame noemit neither; do
  for m in "" "--with-masks"; do
    if [ -z "$v" ]; then L=""; else L="build_variants/$v.so"; fi
    MRTS_CUDA_LIB=$L $B $m 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('variant=[$v] masks=[$m] kernel_ms %.3f' % d['roofline']['mean_launch_ms'])"
  done
done
