#!/bin/bash
for v in "$@" "$@"; do
  echo "== $v"
  MRTS_CUDA_LIB=build_variants/$v.so python bench.py --workload scripted --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('scripted', d['value'], d['ms_per_step'])"
done
