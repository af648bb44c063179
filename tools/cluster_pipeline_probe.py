#!/usr/bin/env python3
"""developer probe (torchrun, one rank per GPU): ONE cluster's first-fit pass with the node list range-sharded
over the ranks and pod blocks pipelined through them (distributed.cluster_first_fit), checked bit for bit
against the single-GPU pass on the same inputs and timed next to it (CUDA events, max over ranks).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
      tools/cluster_pipeline_probe.py --P 300000 --N 100000 --D 8 --T 8
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kubernetes_acs_engine_autoscaler_b200 import distributed as D  # noqa: E402
from kubernetes_acs_engine_autoscaler_b200 import synthetic as syn  # noqa: E402
from kubernetes_acs_engine_autoscaler_b200.engine import Engine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--P", type=int, default=300000)
    ap.add_argument("--N", type=int, default=100000)
    ap.add_argument("--D", type=int, default=8)
    ap.add_argument("--T", type=int, default=8)
    ap.add_argument("--blocks", type=str, default="8,16,32")
    ap.add_argument("--reps", type=int, default=2)
    args = ap.parse_args()
    rank, world, local = D.init()
    eng = Engine(local)
    c = syn.make_cluster(args.P, args.N, args.D, args.T, seed=20260924)
    used0 = syn.initial_used(c)
    f64, i32 = torch.float64, torch.int32
    d_req = eng.dev(c["req"], f64)
    d_cap = eng.dev(c["cap_type"], f64)
    mask, _ = eng.feasible_mask(d_req, eng.dev(c["unit_all"], f64))
    feas = torch.nonzero(mask).flatten().to(i32)

    def timed(fn):
        best, out = 1e30, None
        for _ in range(args.reps):
            D.barrier()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            out = fn()
            b.record()
            torch.cuda.synchronize()
            t = torch.tensor([a.elapsed_time(b)], device=eng.device)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = min(best, float(t.item()))
        return best, out

    # single GPU, whole node list (every rank does it: the reference result for the comparison)
    d_type = eng.dev(c["node_type"], i32)
    used_full = eng.dev(used0, f64)

    def single():
        used_full.copy_(eng.dev(used0, f64))
        return eng.first_fit_nodes(d_req, feas, d_cap, d_type, used_full)
    t1, (placed1, dec1) = timed(single)
    if rank == 0:
        print("single GPU: %.2f ms, %d listed pods, %d placed, %.3e decisions (%.1f G/s)" % (
            t1, feas.numel(), int((placed1 >= 0).sum()), int(dec1.item()), int(dec1.item()) / t1 / 1e6), flush=True)

    N = c["N"]
    lo, hi = (N * rank) // world, (N * (rank + 1)) // world
    d_type_l = eng.dev(c["node_type"][lo:hi], i32)
    for nb in [int(x) for x in args.blocks.split(",")]:
        used_l = eng.dev(used0[lo:hi], f64)

        def sharded():
            used_l.copy_(eng.dev(used0[lo:hi], f64))
            return D.cluster_first_fit(eng, d_req, feas, d_cap, d_type_l, used_l, lo, n_blocks=nb)
        tn, (placedn, decn) = timed(sharded)
        same = bool(torch.equal(placedn.cpu(), placed1.cpu().to(torch.int32))) and int(decn.item()) == int(dec1.item())
        same_used = bool(torch.equal(used_l.cpu().view(torch.int64), used_full[lo:hi].cpu().view(torch.int64)))
        ok = torch.tensor([int(same and same_used)], device=eng.device)
        if world > 1:
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if rank == 0:
            print("%d GPUs, %3d blocks: %.2f ms (x%.2f vs single), bit-exact placements/used/decisions: %s" % (
                world, nb, tn, t1 / tn, bool(ok.item())), flush=True)
    D.shutdown()


if __name__ == "__main__":
    main()
