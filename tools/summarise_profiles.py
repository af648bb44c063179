#!/usr/bin/env python3
"""turn the raw ncu exports in gpurun_out/ into the committed summaries under profiles/ (per round)."""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
src = os.path.join(ROOT, "gpurun_out")
dst = os.path.join(ROOT, "profiles")
os.makedirs(dst, exist_ok=True)

# ---- launch list ------------------------------------------------------------------------------
rows = [r for r in csv.reader(open(os.path.join(src, tag + "_launches.csv"))) if len(r) > 5]
h = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
H, data = rows[h], rows[h + 1:]
ik, iv, iu = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
agg = collections.OrderedDict()
for r in data:
    name = r[ik].split("(")[0]
    v = float(r[iv].replace(",", ""))
    v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r[iu], 1.0)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
with open(os.path.join(dst, tag + "_launches_summary.csv"), "w") as f:
    f.write("kernel,launches,total_ms,share_pct\n")
    for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write('"%s",%d,%.4f,%.2f\n' % (n, a[0], a[1] / 1e6, 100 * a[1] / tot))

# ---- full capture of the pipeline kernel ------------------------------------------------------
rows = list(csv.reader(open(os.path.join(src, tag + "_pipeline_raw.csv"))))
H, units, data = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "smsp__inst_executed.sum",
        "lts__t_bytes.sum", "sm__cycles_elapsed.max"]
traffic = 0.0
with open(os.path.join(dst, tag + "_pipeline_ncu.txt"), "w") as f:
    f.write("ncu --set full --clock-control none, firstfit_pipeline_kernel launches of one bench.py step (c2)\n")
    for d in data:
        f.write("----\n")
        for w in want:
            if w in H:
                i = H.index(w)
                f.write("%-66s %s %s\n" % (w, d[i], units[i]))
        for w in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = H.index(w)
            v = float(d[i].replace(",", ""))
            traffic += v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
with open(os.path.join(dst, tag + "_pipeline_traffic.json"), "w") as f:
    json.dump({"dram_bytes_per_launch": traffic, "what": "dram__bytes_read.sum + dram__bytes_write.sum summed over the "
               "nodes and the bins firstfit_pipeline_kernel launch of one c2 step", "launches": len(data)}, f)
for name in (tag + "_bench.json", tag + "_bench_reference.json", tag + "_gpu.csv"):
    p = os.path.join(src, name)
    if os.path.exists(p):
        with open(p) as g, open(os.path.join(dst, name), "w") as f:
            f.write(g.read())
print("profiles written for", tag, "traffic bytes", traffic)
