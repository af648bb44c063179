#!/bin/bash
# round 2, GPU call E: bulk-copy K1/K6 (parity, ncu time + DRAM bytes at c5, D = 4 and 8, gather form beside it), bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02g_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02g_tests.log
for D in 4 8; do
  timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:node_stream -c 12 \
      --csv --log-file gpurun_out/r02g_idle_D${D}.csv python tools/perf_idle.py --D $D --T $((D==4?1:8)) > gpurun_out/r02g_idle_D${D}.log 2>&1
  ACSFIT_IDLE_GATHER=1 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:node_stream -c 12 \
      --csv --log-file gpurun_out/r02g_idle_gather_D${D}.csv python tools/perf_idle.py --D $D --T $((D==4?1:8)) > gpurun_out/r02g_idle_gather_D${D}.log 2>&1
done
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err
tail -4 gpurun_out/r02g_tests.log
for f in gpurun_out/r02g_idle_D4.csv gpurun_out/r02g_idle_gather_D4.csv gpurun_out/r02g_idle_D8.csv gpurun_out/r02g_idle_gather_D8.csv; do echo $f; python - "$f" <<'PY'
import csv,sys,collections
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>5]
h=[i for i,r in enumerate(rows) if r[0]=="ID"][0]
H=rows[h]; ik,im,iv,iu=H.index("Kernel Name"),H.index("Metric Name"),H.index("Metric Value"),H.index("Metric Unit")
agg=collections.OrderedDict()
for r in rows[h+1:]:
    agg.setdefault((r[0],r[ik][:60]),{})[r[im]]=(r[iv],r[iu])
for (i,k),m in list(agg.items())[-8:]:
    print(i,k,m)
PY
done
tail -c 800 gpurun_out/r02g_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02g_bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['ms_per_step'])
    c3=d['configs']['c3']; print('c3', c3['ms_per_step'], c3['value'], c3['e2e']['ms_per_step'])
    print('c5', {k:(v['node_states']['ms'], v['node_states']['roofline']['frac'], v['occupancy']['ms'], v['occupancy']['roofline']['frac']) for k,v in d['configs']['c5'].items() if k.startswith('D')})
except Exception as e: print('ERR',e)
PY
