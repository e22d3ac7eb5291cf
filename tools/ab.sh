for v in base idlelist base idlelist; do
  echo "== $v"
  MRTS_CUDA_LIB=build_variants/$v.so python bench.py --steps 10 --warmup 15 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('selfplay', d['value'], d['ms_per_step'], d['clocks'])"
done
for v in base idlelist; do
  echo "== $v"
  MRTS_CUDA_LIB=build_variants/$v.so python bench.py --workload rollout --observer -1 --steps 5 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('rollout', d['value'], d['ms_per_step'])"
  MRTS_CUDA_LIB=build_variants/$v.so python bench.py --workload obs --steps 5 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('obs', d['value'], d['ms_per_step'])"
done
