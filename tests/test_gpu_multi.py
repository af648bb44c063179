"""`-m gpu`: cluster mode (include/acsfit.h) -- ONE cluster whose node / bin axis is split over the ranks, the
stage pipeline continued from GPU to GPU through NVLink peer memory inside the kernel -- must reproduce the
single-GPU tick bit for bit: placements, node state (float64 bit patterns), pool sizes, accounted pools, and
the credited decision count.

Two launch shapes:
* two ranks on two GPUs, handles exchanged over NCCL (skipped with fewer than two GPUs);
* "loopback": two ranks (two processes) on ONE GPU, handles exchanged over gloo.  The kernels of the two
  processes are time-sliced by the driver, so this is slow per operation but exercises the identical code
  (IPC mapping, system-scope flags, upstream polling, merges) on a single-GPU box."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

CASES = [
    # P, N, D, T, seed, knobs
    (6000, 700, 4, 2, 7, {}),
    (20000, 3000, 8, 8, 11, {"cluster_blocks": 3}),     # pod blocks forced (the multi-wave schedule)
    (20000, 3000, 4, 1, 12, {"ranks": 0}),               # float64 scan form
    (60000, 6000, 4, 2, 4242, {}),
    (20000, 3000, 8, 8, 13, {"stream": 1}),                # barrier-free streaming form of the pipeline kernel
    (60000, 6000, 4, 2, 4243, {"stream": 1, "cluster_blocks": 2}),
]


def _worker(rank, world, port, out_dir, backend, devices):
    os.environ.update({"RANK": str(rank), "WORLD_SIZE": str(world), "LOCAL_RANK": str(devices[rank]),
                       "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port)})
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from kubernetes_acs_engine_autoscaler_b200 import synthetic as syn
    from kubernetes_acs_engine_autoscaler_b200.engine import Engine
    dev = devices[rank]
    torch.cuda.set_device(dev)
    if backend == "nccl":
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
    else:
        dist.init_process_group("gloo")
    solo = Engine(dev, watchdog_ms=20000)          # the single-GPU answer
    eng = Engine(dev, watchdog_ms=20000)
    eng.cluster_connect(max_pods=70000, max_nodes=8000, max_dims=8)
    eng.configure(watchdog_ms=30000)
    failures = []
    f64, i32 = torch.float64, torch.int32
    for (P, N, D, T, seed, knobs) in CASES:
        c = syn.make_cluster(P, N, D, T, seed=seed, over_provision=1)
        used0 = syn.initial_used(c)
        for e in (solo, eng):
            e.set_knob("ranks", knobs.get("ranks", 1))
            e.set_knob("stream", knobs.get("stream", 0))
        eng.set_knob("cluster_blocks", knobs.get("cluster_blocks", 0))
        args = (c["unit_all"], c["unit_ordered"], c["pool_actual"], c["pool_max"], c["pool_ignored"], c["over_provision"])
        u1, u2 = solo.dev(used0, f64), eng.dev(used0, f64)
        r1 = solo.scale_up(solo.dev(c["req"], f64), *args, solo.dev(c["cap_type"], f64), solo.dev(c["node_type"], i32), u1)
        r2 = eng.scale_up(eng.dev(c["req"], f64), *args, eng.dev(c["cap_type"], f64), eng.dev(c["node_type"], i32), u2)
        torch.cuda.synchronize()
        same = (torch.equal(r1["placed"], r2["placed"]) and torch.equal(r1["acc_pool"], r2["acc_pool"])
                and torch.equal(u1.view(torch.int64), u2.view(torch.int64))
                and np.array_equal(r1["new_size"], r2["new_size"]) and np.array_equal(r1["bins_opened"], r2["bins_opened"])
                and r1["decisions"] == r2["decisions"] and r1["n_pending"] == r2["n_pending"]
                and r1["num_unaccounted"] == r2["num_unaccounted"])
        nontrivial = int((r1["placed"] >= 0).sum()) > 50 and r1["n_pending"] > 50
        if not (same and nontrivial):
            failures.append((P, N, D, T, seed, knobs, same, nontrivial))
        # the host-buffer (plugin) call in cluster mode
        uh = used0.copy()
        r3 = eng.scale_up_host(c["req"], *args, c["cap_type"], c["node_type"], uh)
        if not (np.array_equal(r3["placed"], r1["placed"].cpu().numpy()) and r3["decisions"] == r1["decisions"]
                and np.array_equal(uh.view(np.uint64), u1.cpu().numpy().view(np.uint64))):
            failures.append(("host", P, N, D, T, seed))
    np.save(os.path.join(out_dir, "ok%d.npy" % rank), np.array([0 if failures else 1]))
    if failures:
        sys.stderr.write("rank %d cluster-mode mismatches: %r\n" % (rank, failures))
    dist.barrier()
    dist.destroy_process_group()


def _run(tmp_path, backend, devices):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path), backend, devices), nprocs=2, join=True)
    for rank in range(2):
        assert int(np.load(os.path.join(str(tmp_path), "ok%d.npy" % rank))[0]) == 1


def test_two_gpu_cluster_mode_is_exact(tmp_path):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    _run(tmp_path, "nccl", [0, 1])


def test_loopback_cluster_mode_is_exact(tmp_path):
    """two ranks time-sliced on one GPU: the cluster-mode code path on a single-GPU box."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    _run(tmp_path, "gloo", [0, 0])
