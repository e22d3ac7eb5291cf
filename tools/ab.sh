#!/bin/bash
# A/B of library variants under build_variants/ on the headline benchmark (and the rollout / observation workloads)
for v in "$@" "$@"; do
  echo "== $v"
  MRTS_CUDA_LIB=build_variants/$v.so python bench.py --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('selfplay', d['value'], d['ms_per_step'])"
done
for v in "$@"; do
  echo "== $v"
  MRTS_CUDA_LIB=build_variants/$v.so python bench.py --workload rollout --observer -1 --steps 5 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('rollout', d['value'], d['ms_per_step'])"
  MRTS_CUDA_LIB=build_variants/$v.so python bench.py --map 8x8/basesWorkers8x8 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('8x8', d['value'], d['ms_per_step'])"
done
