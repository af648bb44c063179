#!/bin/bash
# round 2, GPU call C (2 GPUs): cluster mode over real NVLink peers: exactness test, c3 strong scaling at N = 2
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02c_topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r02c_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c_tests.log
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "timeout_falls_back" >> gpurun_out/r02c_tests.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/r02c_bench_n2.json 2> gpurun_out/r02c_bench_n2.err
tail -8 gpurun_out/r02c_tests.log
tail -c 2500 gpurun_out/r02c_bench_n2.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02c_bench_n2.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d.get('strong_scaling'), d['config']['parallelism'], d.get('fleet'), d.get('e2e'))
except Exception as e: print('ERR',e)
PY
