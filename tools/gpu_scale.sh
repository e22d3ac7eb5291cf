#!/bin/bash
# usage: tools/gpu_scale.sh N LABEL  -- bench.py (headline + secondaries) on N GPUs of this box through torchrun
N=$1; L=$2
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${L}_n$N.json 2> gpurun_out/${L}_n$N.err
echo "rc=$?"; tail -3 gpurun_out/${L}_n$N.err
python - $N $L <<'PY'
import json,sys
N,L=sys.argv[1],sys.argv[2]
d=json.loads(open('gpurun_out/%s_n%s.json'%(L,N)).read().strip().splitlines()[-1])
print('N=%s headline %.4g e2e %.4g stats %s'%(N,d['value'],d['e2e']['value'],{k:d['stats'][k] for k in ('wins_p0','wins_p1','draws','games_finished')}))
for k,v in (d.get('secondary') or {}).items(): print(' ',k,v.get('value'),v.get('error'))
PY
