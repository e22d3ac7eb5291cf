#!/bin/bash
# A/B of library variants under build_variants/ on the headline benchmark (twice each), then the 8x8 and rollout workloads
for v in "$@" "$@"; do
  MRTS_CUDA_LIB=build_variants/$v.so python bench.py --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v selfplay', d['value'], d['ms_per_step'])"
done
for v in "$@"; do
  MRTS_CUDA_LIB=build_variants/$v.so python bench.py --workload rollout --observer -1 --steps 5 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v rollout', d['value'], d['ms_per_step'])"
  MRTS_CUDA_LIB=build_variants/$v.so python bench.py --map 8x8/basesWorkers8x8 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v 8x8', d['value'], d['ms_per_step'])"
done
