#!/bin/bash
# round 2, GPU call F: ncu --set full of the K6 / K1 streaming kernels (bulk form), with source counters
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:node_stream_bulk -s 2 -c 4 -o gpurun_out/r02f_idle \
    python tools/perf_idle.py --D 4 --T 1 > gpurun_out/r02f_idle.log 2>&1
ncu -i gpurun_out/r02f_idle.ncu-rep --page details --csv > gpurun_out/r02f_idle_details.csv 2>/dev/null
ncu -i gpurun_out/r02f_idle.ncu-rep --page raw --csv > gpurun_out/r02f_idle_raw.csv 2>/dev/null
ncu -i gpurun_out/r02f_idle.ncu-rep --page source --csv --kernel-name regex:node_stream_bulk --launch-skip 0 --launch-count 1 > gpurun_out/r02f_idle_source.csv 2>/dev/null
ls -la gpurun_out/r02f*
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r02f_idle_raw.csv')))
H,units=rows[0],rows[1]
want=["Kernel Name","gpu__time_duration.sum","launch__registers_per_thread","launch__grid_size","launch__occupancy_limit_registers","launch__occupancy_limit_shared_mem","launch__waves_per_multiprocessor","sm__warps_active.avg.pct_of_peak_sustained_active","smsp__issue_active.avg.pct_of_peak_sustained_active","dram__throughput.avg.pct_of_peak_sustained_elapsed","l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum","smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio","smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio","smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio","smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio","smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio","smsp__average_warps_issue_stalled_wait_per_issue_active.ratio","smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio","smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio","smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio","smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio","smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio","smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio","smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio","smsp__average_warps_issue_stalled_membar_per_issue_active.ratio","smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio"]
for d in rows[2:]:
    print('----')
    for w in want:
        if w in H: print(w, d[H.index(w)], units[H.index(w)])
PY
