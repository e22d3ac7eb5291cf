#!/bin/bash
# one point of the 1/2/4/8 series: the headline benchmark on N GPUs of this box, launched the way the driver does
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-cpu-baseline > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
echo "rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/scale_n$N.json").read().strip().splitlines()[-1])
print($N, d["value"], d["ms_per_step"], d["e2e"] and d["e2e"]["value"], d["roofline"]["frac"])
PY
