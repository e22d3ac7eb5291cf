#!/bin/bash
# CRush_V1 on the device: the scripted differential tests (8 games x 3000 cycles per pairing, 2048-game wide runs)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_engine_parity.py -m gpu -x -q -k "scripted or every_reference_map or partially" 2>&1 | tail -6
timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-secondary --steps 10 --warmup 3 --workload scripted 2>/dev/null | tail -1 | cut -c1-400
