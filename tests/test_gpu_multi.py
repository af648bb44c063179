"""`-m gpu`, needs two visible GPUs (skipped otherwise): the exact one-cluster mode of `distributed.py` -- node
ranges per rank, pod blocks pipelined through the ranks over NCCL point-to-point -- must reproduce the
single-GPU placements, node state (bit patterns) and credited decision count."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out_dir):
    os.environ.update({"RANK": str(rank), "WORLD_SIZE": str(world), "LOCAL_RANK": str(rank),
                       "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port)})
    sys.path.insert(0, ROOT)
    from kubernetes_acs_engine_autoscaler_b200 import distributed as D
    from kubernetes_acs_engine_autoscaler_b200 import synthetic as syn
    from kubernetes_acs_engine_autoscaler_b200.engine import Engine
    _, _, local = D.init("nccl")
    eng = Engine(local)
    c = syn.make_cluster(60000, 6000, 4, 2, seed=4242)
    used0 = syn.initial_used(c)
    f64, i32 = torch.float64, torch.int32
    req = eng.dev(c["req"], f64)
    cap = eng.dev(c["cap_type"], f64)
    mask, _ = eng.feasible_mask(req, eng.dev(c["unit_all"], f64))
    feas = torch.nonzero(mask).flatten().to(i32)
    used_full = eng.dev(used0, f64)
    placed1, dec1 = eng.first_fit_nodes(req, feas, cap, eng.dev(c["node_type"], i32), used_full)
    N = c["N"]
    lo, hi = (N * rank) // world, (N * (rank + 1)) // world
    used_l = eng.dev(used0[lo:hi], f64)
    placed, dec = D.cluster_first_fit(eng, req, feas, cap, eng.dev(c["node_type"][lo:hi], i32), used_l, lo, n_blocks=5)
    ok = (torch.equal(placed.cpu(), placed1.cpu().to(torch.int32)) and int(dec.item()) == int(dec1.item())
          and torch.equal(used_l.cpu().view(torch.int64), used_full[lo:hi].cpu().view(torch.int64))
          and int((placed1 >= 0).sum()) > 100 and int((placed1 < 0).sum()) > 100)
    np.save(os.path.join(out_dir, "ok%d.npy" % rank), np.array([int(ok)]))
    D.barrier()
    D.shutdown()


def test_two_gpu_cluster_pipeline_is_exact(tmp_path):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for rank in range(2):
        assert int(np.load(os.path.join(str(tmp_path), "ok%d.npy" % rank))[0]) == 1
