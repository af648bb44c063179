#!/usr/bin/env python3
"""developer probe: idle-node scan (K6 node_states + K6b maintain_actions + K1 occupancy) at BASELINE
config 5 scale; reports achieved HBM GB/s against the algorithmic bytes of SURVEY 8(d)."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kubernetes_acs_engine_autoscaler_b200 import synthetic as syn  # noqa: E402
from kubernetes_acs_engine_autoscaler_b200.engine import Engine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=1000000)
    ap.add_argument("--D", type=int, default=4)
    ap.add_argument("--T", type=int, default=1)
    ap.add_argument("--S", type=int, default=8)
    args = ap.parse_args()
    eng = Engine(0)
    c = syn.make_idle_cluster(args.N, args.D, args.T, seed=5)
    R = c["req_run"].shape[0]
    i64, i32, f64, u8 = torch.int64, torch.int32, torch.float64, torch.uint8
    ap_gather = os.environ.get("ACSFIT_IDLE_GATHER") == "1"   # 1: pass the index list (gather kernel) instead of NULL
    d = {k: eng.dev(c[k], t) for k, t in (("row_ptr", i64), ("run_idx", i32), ("req_run", f64), ("flags_run", u8),
                                          ("cap_type", f64), ("node_type", i32), ("node_flags", u8), ("node_age", i64),
                                          ("node_pool", i32))}
    if not ap_gather:
        d["run_idx"] = None  # contiguous table: the bulk-copy streaming kernel
    thr = np.array([60, 300, 900, 1800, 3600, 7200, 21600, 86400][:args.S], dtype=np.int64)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=eng.device)

    def timed(fn, reps=5):
        best = 1e30
        for _ in range(reps):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            out = fn()
            b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        return best, out

    D, N, S = args.D, args.N, args.S
    t, st = timed(lambda: eng.node_states(d["row_ptr"], d["run_idx"], d["req_run"], d["flags_run"], d["cap_type"],
                                          d["node_type"], d["node_flags"], d["node_age"], False, thr))
    bytes_states = R * (8 * D + 1 + (4 if ap_gather else 0)) + N * (8 + 4 + 1 + 8) + S * N
    print("node_states: N=%d R=%d D=%d S=%d: %.3f ms, %.1f GB/s algorithmic (%.1f MB)" % (N, R, D, S, t, bytes_states / t / 1e6, bytes_states / 1e6))
    budget = c["pool_actual"].astype(np.int64) - 1
    for dry in (True, False):
        s0 = st[3 if S > 3 else 0].clone()
        t, _ = timed(lambda: eng.maintain_actions(s0, d["node_pool"], budget, np.ones(args.T, np.uint8), dry))
        print("maintain_actions dry_run=%s: %.3f ms (%.1f GB/s over %d B/node)" % (dry, t, N * 6 / t / 1e6, 6))
    used = torch.zeros((N, D), dtype=f64, device=eng.device)
    t, _ = timed(lambda: eng.occupancy(d["row_ptr"], d["run_idx"], d["req_run"], used.zero_()))
    bytes_occ = R * (8 * D + (4 if ap_gather else 0)) + N * (8 + 16 * D)
    print("occupancy: %.3f ms, %.1f GB/s algorithmic" % (t, bytes_occ / t / 1e6))


if __name__ == "__main__":
    main()
