#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02m_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02m_tests.log
tail -4 gpurun_out/r02m_tests.log
timeout 600 python tools/stress_parity.py > gpurun_out/r02m_stress.log 2>&1; tail -2 gpurun_out/r02m_stress.log
timeout 300 python tools/perf_probe.py 2>&1 | grep "min_stages\|fulfill" | sed "s/^/c2: /"
timeout 600 python tools/perf_probe.py --P 1000000 --N 100000 --D 8 --T 8 2>&1 | grep "min_stages\|fulfill" | sed "s/^/c3: /"
timeout 900 python bench.py --steps 10 --warmup 3 2>gpurun_out/r02m_bench.err > gpurun_out/r02m_bench.json
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02m_bench.json').read().strip().splitlines()[-1])
    print('c2 ms', d['ms_per_step'], 'kernel', d['roofline']['kernel_ms_per_step'], d['pipeline'], 'e2e', d['e2e']['ms_per_step'], 'c3 ms', d['configs']['c3']['ms_per_step'], 'python_surface', d['e2e'].get('python_surface',{}).get('seconds_per_tick'), 'traffic', d['roofline']['traffic'], d['roofline']['physical'].get('chain_ns_per_placement'))
except Exception as e: print('ERR',e); print(open('gpurun_out/r02m_bench.err').read()[-1500:])
PY
