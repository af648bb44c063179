#!/bin/bash
# round 2 profile collection (one GPU, through gpurun): launch list + full captures of the dominant kernels + the
# streaming kernels + the bench lines.  Raw exports go to gpurun_out/, tools/summarise_r02.py condenses them into profiles/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/r02_gpu.csv
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-sub > gpurun_out/r02_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:firstfit_pipeline -s 4 -c 2 \
    -o gpurun_out/r02_pipeline_c2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sub > gpurun_out/r02_full_c2.log 2>&1
ncu -i gpurun_out/r02_pipeline_c2.ncu-rep --page raw --csv > gpurun_out/r02_pipeline_c2_raw.csv 2>/dev/null
ncu --set full --clock-control none -k regex:firstfit_pipeline -s 7 -c 7 \
    -o gpurun_out/r02_pipeline_c3 python bench.py --config c3 --steps 1 --warmup 1 --no-cpu-baseline --no-sub > gpurun_out/r02_full_c3.log 2>&1
ncu -i gpurun_out/r02_pipeline_c3.ncu-rep --page raw --csv > gpurun_out/r02_pipeline_c3_raw.csv 2>/dev/null
rm -f gpurun_out/r02_pipeline_c3.ncu-rep   # (41 MB: the raw CSV export is what travels back; gpurun merges at most 64 MiB)
for D in 4 8; do
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:node_stream -c 10 \
      --csv --log-file gpurun_out/r02_idle_D${D}.csv python tools/perf_idle.py --D $D --T $((D==4?1:8)) > gpurun_out/r02_idle_D${D}.log 2>&1
done
cuobjdump -sass kubernetes_acs_engine_autoscaler_b200/libacsfit.so | grep -c "UBLKCP" > gpurun_out/r02_sass_ublkcp_count.txt
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
tail -c 400 gpurun_out/r02_bench.json; ls -la gpurun_out/r02_pipeline_c*.ncu-rep
