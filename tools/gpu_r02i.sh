#!/bin/bash
# round 2, GPU call I (N GPUs): full parity suite (incl. the N-GPU cluster test), then the scaling bench with per-pass trace
N=${1:-2}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02i_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02i_tests.log
tail -3 gpurun_out/r02i_tests.log
ACSFIT_TRACE=1 timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 \
    bench.py --gpus $N --steps 3 --warmup 2 --no-fleet > gpurun_out/r02i_scale_n$N.json 2> gpurun_out/r02i_scale_n$N.err
grep "acsfit" gpurun_out/r02i_scale_n$N.err | grep -v "r[1-9]\]" | tail -30
python - $N <<'PY'
import json,sys
N=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/r02i_scale_n%s.json'%N).read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d.get('strong_scaling'), d['config']['parallelism'])
except Exception as e: print('ERR',e)
PY
