#!/bin/bash
# round 2, GPU call D: parity after the resolver prefilter / early alive load / fence change; c3 + c2 probes; bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02d_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d_tests.log
timeout 600 python tools/stress_parity.py > gpurun_out/r02d_stress.log 2>&1; echo "stress rc=$?" >> gpurun_out/r02d_stress.log
timeout 600 python tools/perf_probe.py --P 1000000 --N 100000 --D 8 --T 8 --prof --trace-stage 100 > gpurun_out/r02d_probe_c3.log 2>&1
timeout 300 python tools/perf_probe.py --prof --trace-stage 20 > gpurun_out/r02d_probe_c2.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err
tail -4 gpurun_out/r02d_tests.log; tail -3 gpurun_out/r02d_stress.log
grep -v "^     \|^        " gpurun_out/r02d_probe_c3.log | head -22
grep -A12 "trace of stage 100" gpurun_out/r02d_probe_c3.log | head -14
grep "min_stages\|fulfill_pending" gpurun_out/r02d_probe_c2.log
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02d_bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['ms_per_step'])
    c3=d['configs']['c3']; print('c3', c3['ms_per_step'], c3['value'], c3['e2e']['ms_per_step'])
    print('c5', {k:(v['node_states']['ms'], v['occupancy']['ms']) for k,v in d['configs']['c5'].items() if k.startswith('D')})
except Exception as e: print('ERR',e)
PY
