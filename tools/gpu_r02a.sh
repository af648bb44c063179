#!/bin/bash
# round 2, GPU call A: parity suite (both scan forms, loopback cluster mode), bench c2 / c3 with either scan
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02a_gpus.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r02a_bench_c2.json 2> gpurun_out/r02a_bench_c2.err
ACSFIT_RANKS=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_bench_c2_f64.json 2> gpurun_out/r02a_bench_c2_f64.err
timeout 300 python bench.py --config c3 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r02a_bench_c3.json 2> gpurun_out/r02a_bench_c3.err
ACSFIT_RANKS=0 timeout 300 python bench.py --config c3 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r02a_bench_c3_f64.json 2> gpurun_out/r02a_bench_c3_f64.err
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -x -q -k "first_fit_nodes_matches_oracle or fulfill" > gpurun_out/r02a_sanitizer.log 2>&1
tail -5 gpurun_out/r02a_tests.log
for f in gpurun_out/r02a_bench_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'])
except Exception as e: print('ERR', e)
"; done
tail -3 gpurun_out/r02a_sanitizer.log
