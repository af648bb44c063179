#!/bin/bash
# round 2, GPU call H: geometry sweep of the bulk-copy K1/K6 (CUDA-event timing from perf_idle, L2 flushed)
mkdir -p gpurun_out
for D in 4 8; do for C in 0 1 2 3 4; do
  echo "== D=$D cfg=$C" >> gpurun_out/r02h_sweep.log
  ACSFIT_BULK_CFG=$C timeout 200 python tools/perf_idle.py --D $D --T $((D==4?1:8)) 2>&1 | grep "node_states\|occupancy" >> gpurun_out/r02h_sweep.log
done; done
cat gpurun_out/r02h_sweep.log
