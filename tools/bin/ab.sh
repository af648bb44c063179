L=kubernetes_acs_engine_autoscaler_b200/libacsfit.so
for rep in 1 2 3; do for v in base split; do cp tools/bin/lib_$v.so $L; echo -n "$v: "; python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; b=json.loads(sys.stdin.read()); print(round(b['ms_per_step'],3), round(b['value']/1e9,1))"; done; done
cp tools/bin/lib_split.so $L; timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -1
