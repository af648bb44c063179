#!/usr/bin/env python3
"""turn the raw ncu exports of tools/collect_profiles_r02.sh (gpurun_out/r02_*) into the committed summaries under
profiles/: launch-list shares, the counters of the full captures, DRAM traffic per kernel (profiles/r02_traffic.json,
read by bench.py as roofline.traffic), the streaming kernels' time / bytes, and copies of the bench lines."""
import collections
import csv
import json
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src, dst = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(dst, exist_ok=True)
UNIT = {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9, "usecond": 1e3, "nsecond": 1.0, "msecond": 1e6, "second": 1e9}
BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def short(name):
    name = name.replace("void ", "").replace("acsfit::", "").replace("<unnamed>::", "")
    return name.split("(")[0]


# ---- launch list ------------------------------------------------------------------------------
rows = [r for r in csv.reader(open(os.path.join(src, "r02_launches.csv"))) if len(r) > 5]
h = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
H, data = rows[h], rows[h + 1:]
ik, iv, iu = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
agg = collections.OrderedDict()
for r in data:
    a = agg.setdefault(short(r[ik]), [0, 0.0])
    a[0] += 1
    a[1] += float(r[iv].replace(",", "")) * UNIT.get(r[iu], 1.0)
tot = sum(a[1] for a in agg.values())
with open(os.path.join(dst, "r02_launches_summary.csv"), "w") as f:
    f.write("kernel,launches,total_ms,share_pct\n")
    for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write('"%s",%d,%.4f,%.2f\n' % (n, a[0], a[1] / 1e6, 100 * a[1] / tot))

# ---- full captures of the pipeline kernel -----------------------------------------------------
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "smsp__inst_executed.sum", "lts__t_bytes.sum", "sm__cycles_elapsed.max"]
traffic = {}
for cfg in ("c2", "c3"):
    rows = list(csv.reader(open(os.path.join(src, "r02_pipeline_%s_raw.csv" % cfg))))
    H, units, data = rows[0], rows[1], rows[2:]
    total, counted = 0.0, 0
    with open(os.path.join(dst, "r02_pipeline_%s_ncu.txt" % cfg), "w") as f:
        f.write("ncu --set full --clock-control none: the firstfit_pipeline_kernel launches of one bench.py step (%s)\n" % cfg)
        for d in data:
            f.write("----\n")
            for w in want:
                if w in H:
                    i = H.index(w)
                    f.write("%-66s %s %s\n" % (w, short(d[i]) if w == "Kernel Name" else d[i], units[i]))
            vals = [float(d[H.index(w)].replace(",", "")) * BYTES[units[H.index(w)]] for w in ("dram__bytes_read.sum", "dram__bytes_write.sum")]
            if all(v == v for v in vals):  # ncu reports -nan for the 1024-stage bin launches of c3 (multi-pass replay)
                total += sum(vals)
                counted += 1
    if cfg == "c3":  # several steps were captured and the big bin launches have no DRAM counters: report the node pass alone
        ik = H.index("Kernel Name")
        node_vals = []
        for d in data:
            vals = [float(d[H.index(w)].replace(",", "")) * BYTES[units[H.index(w)]] for w in ("dram__bytes_read.sum", "dram__bytes_write.sum")]
            if ", 0, 256" in d[ik] and all(v == v for v in vals):
                node_vals.append(sum(vals))
        total = sum(node_vals) / max(1, len(node_vals))
        traffic["pipeline_c3_what"] = "mean DRAM bytes of the node-pass launch (the dominant kernel of a c3 step)"
    traffic["pipeline_%s" % cfg] = total
    traffic["pipeline_%s_launches" % cfg] = len(data)
    traffic["pipeline_%s_launches_with_dram_counters" % cfg] = counted

# ---- streaming kernels (K1 / K6) --------------------------------------------------------------
idle = {}
for D in (4, 8):
    rows = [r for r in csv.reader(open(os.path.join(src, "r02_idle_D%d.csv" % D))) if len(r) > 5]
    h = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    H = rows[h]
    ik, im, iv, iu = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value"), H.index("Metric Unit")
    per = collections.OrderedDict()
    for r in rows[h + 1:]:
        per.setdefault((r[0], short(r[ik])), {})[r[im]] = float(r[iv].replace(",", "")) * (UNIT.get(r[iu]) or BYTES.get(r[iu]) or 1.0)
    for kind, tag in ((", 1, ", "k6"), (", 0, ", "k1")):
        sel = [m for (i, k), m in per.items() if kind in k]
        if sel:
            t = sorted(m["gpu__time_duration.sum"] for m in sel)[len(sel) // 2]
            b = sorted(m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"] for m in sel)[len(sel) // 2]
            idle["%s_D%d" % (tag, D)] = {"kernel": [k for (i, k), m in per.items() if kind in k][0], "launches": len(sel),
                                         "median_us": t / 1e3, "dram_bytes": b, "dram_GBps": b / t}
            traffic["%s_D%d" % (tag, D)] = b
with open(os.path.join(dst, "r02_idle_ncu.json"), "w") as f:
    json.dump(idle, f, indent=1)
traffic["what"] = ("dram__bytes_read.sum + dram__bytes_write.sum (ncu): pipeline_* summed over the firstfit_pipeline_kernel launches "
                   "of one bench.py step, k1_* / k6_* per launch of the streaming kernels at 1M nodes")
with open(os.path.join(dst, "r02_traffic.json"), "w") as f:
    json.dump(traffic, f, indent=1)
for name in ("r02_bench.json", "r02_bench_reference.json", "r02_gpu.csv", "r02_scale_n4.json", "r02_scale_n8.json",
             "r02_sass_ublkcp_count.txt", "r02b_ubench6.txt", "r02_scale_n2.json"):
    p = os.path.join(src, name)
    if os.path.exists(p):
        shutil.copy(p, os.path.join(dst, name))
print(json.dumps(traffic, indent=1))
print(json.dumps(idle, indent=1))
print(open(os.path.join(dst, "r02_launches_summary.csv")).read())
