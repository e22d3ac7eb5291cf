#!/bin/bash
# 8-GPU and 1-GPU runs of the headline benchmark on the same box (launched the way the driver does)
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
echo "rc=$?"; tail -c 400 gpurun_out/scale_n$N.err
python bench.py --gpus 1 --no-cpu-baseline > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
echo "rc=$?"
python - <<PY
import json
for n in ($N, 1):
    d = json.loads(open("gpurun_out/scale_n%d.json" % n).read().strip().splitlines()[-1])
    print(n, d["value"], d["ms_per_step"], d["e2e"] and d["e2e"]["value"], d["roofline"]["frac"], d["clocks"])
PY
