#!/bin/bash
# round 2, GPU call K: the streaming form of the pipeline kernel: parity (all tests x 3 engine modes), stress, probes, bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02l_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02l_tests.log
tail -6 gpurun_out/r02l_tests.log
ACSFIT_STREAM=1 timeout 600 python tools/stress_parity.py > gpurun_out/r02l_stress.log 2>&1; echo "stress rc=$?" >> gpurun_out/r02l_stress.log; tail -3 gpurun_out/r02l_stress.log
for S in 0 1; do
  ACSFIT_STREAM=$S timeout 300 python tools/perf_probe.py 2>&1 | grep "min_stages\|fulfill" | sed "s/^/stream=$S c2: /"
  ACSFIT_STREAM=$S timeout 600 python tools/perf_probe.py --P 1000000 --N 100000 --D 8 --T 8 2>&1 | grep "min_stages\|fulfill" | sed "s/^/stream=$S c3: /"
  ACSFIT_STREAM=$S timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/r02l_bench_s$S.err > gpurun_out/r02l_bench_s$S.json
  python - $S <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/r02l_bench_s%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
    print('stream=%s'%sys.argv[1], 'c2 ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'c3 ms', d['configs']['c3']['ms_per_step'], 'python_surface', d['e2e'].get('python_surface'))
except Exception as e: print('ERR',e); print(open('gpurun_out/r02l_bench_s%s.err'%sys.argv[1]).read()[-1500:])
PY
done
