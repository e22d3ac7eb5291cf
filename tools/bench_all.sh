#!/bin/bash
# every bench line of DESIGN.md section 7 with the committed library (one B200)
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py > gpurun_out/b_main.json 2> gpurun_out/b_main.err
python bench.py --workload obs > gpurun_out/b_obs.json 2> gpurun_out/b_obs.err
python bench.py --workload scripted > gpurun_out/b_scr.json 2> gpurun_out/b_scr.err
python bench.py --workload rollout > gpurun_out/b_roll.json 2> gpurun_out/b_roll.err
python bench.py --workload rollout --observer -1 > gpurun_out/b_rollfo.json 2> gpurun_out/b_rollfo.err
python bench.py --map 8x8/basesWorkers8x8 > gpurun_out/b_8x8.json 2> gpurun_out/b_8x8.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/b_ref.json 2> gpurun_out/b_ref.err
python - <<PY
import json
for f in ("main","obs","scr","roll","rollfo","8x8","ref"):
    try:
        d=json.loads(open("gpurun_out/b_%s.json"%f).read().strip().splitlines()[-1])
        print(f, "%.4g"%d["value"], "ms/step %.3f"%d["ms_per_step"], "e2e", d.get("e2e") and "%.4g"%d["e2e"]["value"], "frac", d.get("roofline") and d["roofline"].get("frac"), "cpu", d.get("cpu_baseline") and "%.4g"%d["cpu_baseline"]["value"], {k:v for k,v in d.items() if k in ("rollouts_per_s","mean_rollout_cycles")})
    except Exception as e:
        print(f, "FAILED", e)
PY
