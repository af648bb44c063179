#!/bin/bash
# round 2 final single-GPU pass: parity suite, smoke, stress in the alternative schedules, then the profile collection
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_final_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_final_tests.log
tail -3 gpurun_out/r02_final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final_smoke.log 2>&1; tail -1 gpurun_out/r02_final_smoke.log
timeout 600 python tools/stress_parity.py --cases 300 --seed 7 > gpurun_out/r02_final_stress.log 2>&1; tail -1 gpurun_out/r02_final_stress.log
ACSFIT_PRUNE=1 timeout 600 python tools/stress_parity.py --cases 150 --seed 8 >> gpurun_out/r02_final_stress.log 2>&1; tail -1 gpurun_out/r02_final_stress.log
ACSFIT_OVERLAP=0 ACSFIT_RANKS=0 timeout 600 python tools/stress_parity.py --cases 150 --seed 9 >> gpurun_out/r02_final_stress.log 2>&1; tail -1 gpurun_out/r02_final_stress.log
bash tools/collect_profiles_r02.sh > gpurun_out/r02_final_collect.log 2>&1
tail -c 300 gpurun_out/r02_bench.json
