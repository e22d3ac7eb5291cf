#!/bin/bash
# soak: the randomized differential tests of tests/test_engine_parity.py with other seeds (MRTS_SEED_OFFSET), on the GPU
mkdir -p gpurun_out
for off in ${SOAK_OFFSETS:-1000003 2000003 3000017 4000037}; do
  start=$(date +%s)
  MRTS_SEED_OFFSET=$off timeout 1500 python -m pytest tests/test_engine_parity.py -m gpu -x -q > gpurun_out/soak_$off.log 2>&1
  echo "offset $off rc=$? $(( $(date +%s) - start )) s: $(tail -1 gpurun_out/soak_$off.log)"
done
