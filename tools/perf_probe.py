#!/usr/bin/env python3
"""developer probe: per-phase device timings of the tick at a given shape (not a bench value).
--prof needs a library built with the stage profile: python -m kubernetes_acs_engine_autoscaler_b200.build --force --profile
(rebuild without --profile afterwards: the probe costs registers in the placement loop)."""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kubernetes_acs_engine_autoscaler_b200 import synthetic as syn  # noqa: E402
from kubernetes_acs_engine_autoscaler_b200.engine import Engine  # noqa: E402


def ev_time(fn, reps=3):
    best = 1e30
    out = None
    for _ in range(reps):
        a = torch.cuda.Event(enable_timing=True)
        b = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best, out


def unpack_warp(v):
    v = int(v)
    return ((v & 0x7FFF) * 32, ((v >> 15) & 0x7FFF) * 32, (v >> 30) & 0x1FF, (v >> 39) & 0x3F, (v >> 45) & 0x3FF, (v >> 55) & 0x1FF)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--P", type=int, default=100000)
    ap.add_argument("--N", type=int, default=10000)
    ap.add_argument("--D", type=int, default=4)
    ap.add_argument("--T", type=int, default=1)
    ap.add_argument("--stages", type=str, default="0")
    ap.add_argument("--prof", action="store_true")
    ap.add_argument("--trace-stage", type=int, default=100)
    ap.add_argument("--full-nodes", action="store_true", help="every node exactly full: pure scan, no hits")
    args = ap.parse_args()
    eng = Engine(0)
    c = syn.make_cluster(args.P, args.N, args.D, args.T, seed=20260923)
    used0 = syn.initial_used(c)
    if args.full_nodes:
        used0 = c["cap_type"][c["node_type"]].copy()
    f64, i32 = torch.float64, torch.int32
    d_req = eng.dev(c["req"], f64)
    d_cap = eng.dev(c["cap_type"], f64)
    d_type = eng.dev(c["node_type"], i32)
    d_used0 = eng.dev(used0, f64)
    d_unit = eng.dev(c["unit_all"], f64)
    mask, _ = eng.feasible_mask(d_req, d_unit)
    feas = torch.nonzero(mask).flatten().to(i32)
    eng.set_timing(True)
    if args.prof:
        eng.debug_profile(True)
        eng.debug_trace(args.trace_stage)

    def show_prof(tag):
        if not args.prof:
            return
        pr = eng.debug_profile(True).astype(np.float64)
        if not len(pr):
            return
        names = ["wait", "load", "scan", "resolve", "publish", "refresh"]
        tot = pr[:, :6].sum(axis=1)
        print("   [%s] stages %d; per-stage busy cycles: max %.3g mean %.3g" % (tag, len(pr), tot.max(), tot.mean()))
        print("   resolver entries over all stages: %d (tiles with hits: %d)" % (pr[:, 6].sum(), (pr[:, 6] > 0).sum()))
        print("   phase sums over stages (Mcycles): " + ", ".join("%s %.1f" % (n, pr[:, i].sum() / 1e6) for i, n in enumerate(names)))
        tr = eng.debug_trace(args.trace_stage).astype(np.int64)
        act = np.nonzero(tr[:, :6].sum(axis=1))[0]
        if len(act):
            print("   trace of stage %d (tile: wait load scan resolve publish refresh | hits alive):" % args.trace_stage)
            sel = list(act[:6]) + list(act[np.argsort(-tr[act, 3])[:10]])
            for t_ in sel:
                print("     tile %4d: %s | %d %d" % (t_, " ".join("%7d" % v for v in tr[t_, :6]), tr[t_, 6], tr[t_, 7]))
                if t_ < 4096 and tr[t_, 6] > 0:
                    w = tr[4096 + t_]
                    print("        warps (batch cycles/of which loop+bookkeeping, entries seen, batches, loop iterations, placed): " + "  ".join(
                        "w%d %d/%d e%d b%d i%d p%d" % ((i,) + unpack_warp(v)) for i, v in enumerate(w) if v))
            ws = np.array([[unpack_warp(v) for v in tr[4096 + t_]] for t_ in act if t_ < 4096], dtype=np.int64)
            if len(ws):
                tot_w = ws.sum(axis=0)   # [warp, field]
                print("   stage %d totals per warp over %d tiles (batch cycles, loop cycles, entries, batches, iterations, placed):" % (args.trace_stage, len(ws)))
                for i in range(tot_w.shape[0]):
                    if tot_w[i].any():
                        print("     w%d: %s" % (i, " ".join("%8d" % v for v in tot_w[i])))
                print("     all: iterations %d, placed %d, entries %d, resolve cycles %d" % (tot_w[:, 4].sum(), tot_w[:, 5].sum(), tr[act, 6].sum(), tr[act, 3].sum()))
        top = np.argsort(-(tot - pr[:, 0]))[:4]
        for s_ in top:
            print("   stage %4d: " % s_ + ", ".join("%s %.0fk" % (n, pr[s_, i] / 1e3) for i, n in enumerate(names))
                  + ", hits %d tiles %d" % (pr[s_, 6], pr[s_, 7]))
    for ms in [int(x) for x in args.stages.split(",")]:
        eng.configure(min_stages=ms)
        used = d_used0.clone()

        def ff():
            used.copy_(d_used0)
            return eng.first_fit_nodes(d_req, feas, d_cap, d_type, used)
        t, (placed, dec) = ev_time(ff)
        s = eng.pipeline_stats()
        dec = int(dec.item())
        npl = int((placed >= 0).sum().item())
        show_prof("nodes")
        print("min_stages=%d first_fit_nodes: total %.3f ms, pipeline kernel %.3f ms, stages %d tiles %d, decisions %.3e (%.1f G/s kernel), placed %d"
              % (ms, t, s["ms"], s["stages"], s["tiles"], dec, dec / s["ms"] / 1e6, npl))
        pend = feas[placed < 0]
        req_p = d_req[pend.long()].contiguous()

        def fp():
            return eng.fulfill_pending(req_p, req_p.shape[0], c["unit_ordered"], c["pool_actual"], c["pool_max"],
                                       c["pool_ignored"], 0)
        t, r = ev_time(fp)
        s = eng.pipeline_stats()
        show_prof("bins (last pass)")
        print("             fulfill_pending: total %.3f ms, pipeline kernels %.3f ms, stages %d tiles %d, evals %.3e (%.1f G/s kernel), pending %d bins %s"
              % (t, s["ms"], s["stages"], s["tiles"], r["evals"], r["evals"] / max(s["ms"], 1e-9) / 1e6,
                 req_p.shape[0], r["bins_opened"].tolist()))


if __name__ == "__main__":
    main()
