#!/bin/bash
# run on the GPU box through gpurun: launch list + full capture of the dominant kernel (nodes and bins
# instances) for the bench configuration.  Outputs go to gpurun_out/ and are summarised into profiles/.
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r01_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:firstfit_pipeline -s 4 -c 2 \
    -o gpurun_out/r01_pipeline python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r01_full_capture.log 2>&1
ncu -i gpurun_out/r01_pipeline.ncu-rep --page raw --csv > gpurun_out/r01_pipeline_raw.csv 2>/dev/null
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/r01_gpu.csv
python bench.py --steps 10 --warmup 3 > gpurun_out/r01_bench.json 2> gpurun_out/r01_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01_bench_reference.json 2>/dev/null
tail -c 600 gpurun_out/r01_bench.json
# idle-node scan (BASELINE config 5): launch list with DRAM bytes of the streaming kernels
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:node_stream -c 10 \
    --csv --log-file gpurun_out/r01_idle_launches.csv python tools/perf_idle.py > gpurun_out/r01_idle.log 2>&1
