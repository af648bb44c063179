#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/perf_probe.py --stages 0,300,600,1200 > gpurun_out/r02j_probe_c2_stages.log 2>&1
timeout 900 python tools/perf_probe.py --P 1000000 --N 100000 --D 8 --T 8 --stages 0,800,1600 > gpurun_out/r02j_probe_c3_stages.log 2>&1
for S in 0 300 600; do ACSFIT_MIN_STAGES=$S timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sub 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('min_stages=$S', d['ms_per_step'], d['pipeline'])"; done
cat gpurun_out/r02j_probe_c2_stages.log gpurun_out/r02j_probe_c3_stages.log | grep "min_stages\|fulfill"
