#!/usr/bin/env python3
"""Config 5 batch-size sweep: bench.py --workload obs for 1 Ki .. 256 Ki games on the GPUs of this launch.
usage: python tools/sweep_obs.py > profiles/<round>_cfg5_sweep.jsonl"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for games in (1024, 4096, 16384, 65536, 262144):
    steps = 600 if games <= 16384 else (300 if games == 65536 else 100)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "obs", "--games", str(games), "--steps", str(steps),
                          "--no-cpu-baseline", "--prewarm-seconds", "0.5"], capture_output=True, text=True)
    line = out.stdout.strip().splitlines()[-1] if out.stdout.strip() else json.dumps(dict(games=games, error=out.stderr[-400:]))
    d = json.loads(line)
    print(json.dumps(dict(games=games, value=d.get("value"), ms_per_step=d.get("ms_per_step"), achieved_gbs=(d.get("roofline") or {}).get("achieved"),
                          frac=(d.get("roofline") or {}).get("frac"), mean_live_units=(d.get("config") or {}).get("mean_live_units"), error=d.get("error"))))
    sys.stdout.flush()
