#!/bin/bash
# config 3 alone: three runs of the scripted workload + the scripted parity tests
mkdir -p gpurun_out
for i in 1 2 3; do
python bench.py --no-cpu-baseline --no-e2e --no-secondary --steps 10 --warmup 3 --workload scripted 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg3 %.4g game-cycles/s %.2f ms/step'%(d['value'], d['ms_per_step']))"
done
timeout 1200 python -m pytest tests/test_engine_parity.py tests/test_benchmark_shape.py -m gpu -x -q -k "scripted or lightrush or cfg3 or config3 or rush" 2>&1 | tail -3
