#!/bin/bash
# usage: ab_scr.sh "<extra bench flags>" variant...
X="$1"; shift
for v in "$@" "$@"; do
  MRTS_CUDA_LIB=build_variants/$v.so python bench.py --workload scripted --no-cpu-baseline --steps 40 $X 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v $X scripted', d['value'], d['ms_per_step'], d['stats']['game_errors'])"
done
