#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3_pytest.log
tail -3 gpurun_out/s3_pytest.log
B="python bench.py --no-cpu-baseline --no-e2e --no-secondary --steps 40 --warmup 5"
$B --workload obs > gpurun_out/s3_obs.json 2>gpurun_out/s3_obs.err
$B --workload obs --with-masks > gpurun_out/s3_obsm.json 2>gpurun_out/s3_obsm.err
$B --workload obs --no-stagger > gpurun_out/s3_obs_ns.json 2>gpurun_out/s3_obs_ns.err
$B > gpurun_out/s3_self.json 2>gpurun_out/s3_self.err
$B --workload cfg1 > gpurun_out/s3_cfg1.json 2>gpurun_out/s3_cfg1.err
$B --workload rollout --steps 10 > gpurun_out/s3_roll.json 2>gpurun_out/s3_roll.err
$B --workload scripted --steps 10 > gpurun_out/s3_scr.json 2>gpurun_out/s3_scr.err
for f in obs obsm obs_ns self cfg1 roll scr; do python - $f <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/s3_%s.json'%f).read().strip().splitlines()[-1])
    r=d['roofline']
    print(f, '%.4g'%d['value'], 'ms/step %.3f'%d['ms_per_step'], 'kernel_ms %.3f'%r['mean_launch_ms'], r['kernel'], 'frac %.3f'%r['frac'], 'achieved %.0f GB/s'%r['achieved'], 'dram_frac %.3f'%r['dram_frac'])
except Exception as e:
    print(f, 'FAILED', e); print(open('gpurun_out/s3_%s.err'%f).read()[-800:])
PY
done
