#!/bin/bash
# round 2, GPU call B: where does the c3 tick spend its time (phase profile, both scan forms), int issue ubench, new bench line
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench6 tools/ubench6.cu && tools/ubench6 > gpurun_out/r02b_ubench6.txt 2>&1
for R in 1 0; do
  ACSFIT_RANKS=$R timeout 600 python tools/perf_probe.py --P 1000000 --N 100000 --D 8 --T 8 --prof --trace-stage 100 > gpurun_out/r02b_probe_c3_ranks$R.log 2>&1
  ACSFIT_RANKS=$R timeout 600 python tools/perf_probe.py --P 1000000 --N 100000 --D 8 --T 8 --full-nodes > gpurun_out/r02b_probe_c3_full_ranks$R.log 2>&1
done
ACSFIT_RANKS=1 timeout 300 python tools/perf_probe.py --prof --trace-stage 20 > gpurun_out/r02b_probe_c2_ranks1.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err
cat gpurun_out/r02b_ubench6.txt
for R in 1 0; do echo "== ranks=$R"; grep -v "^     \|^        " gpurun_out/r02b_probe_c3_ranks$R.log | tail -12; tail -3 gpurun_out/r02b_probe_c3_full_ranks$R.log; done
tail -c 1500 gpurun_out/r02b_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02b_bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['ms_per_step'], d['roofline']['physical'])
    print(json.dumps(d.get('configs'))[:3000])
    print(json.dumps(d.get('cpu_baseline'))[:1200])
except Exception as e: print('ERR',e)
PY
