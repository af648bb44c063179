#!/bin/bash
# round 2: the driver's scaling command at N GPUs (one cluster, strong scaling), plus the reference arm
N=${1:-4}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo_n$N.txt 2>&1
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 \
    bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_scale_n$N.json 2> gpurun_out/r02_scale_n$N.err
tail -c 1500 gpurun_out/r02_scale_n$N.err
python - $N <<'PY'
import json,sys
N=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/r02_scale_n%s.json'%N).read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d.get('strong_scaling'), d['config']['parallelism'])
    print('fleet', d.get('fleet')); print('e2e', d.get('e2e'))
    print('c4', json.dumps(d.get('configs'))[:2500])
except Exception as e: print('ERR',e)
PY
