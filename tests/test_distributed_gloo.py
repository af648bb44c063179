"""`-m "not gpu"`: the N>1 path (independent cluster shards + one all-reduce of the per-pool integer
counts) with world_size 2 on the gloo backend; the oracle stands in for the GPU engine."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def shard_tick(rank):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle
    from kubernetes_acs_engine_autoscaler_b200 import synthetic as syn
    from test_gpu_parity import oracle_scale_up
    c = syn.make_cluster(3000, 300, 4, 2, seed=100 + rank, over_provision=rank)
    return c, oracle_scale_up(oracle, c, syn.initial_used(c))


def worker(rank, world, port, out_dir):
    os.environ.update({"RANK": str(rank), "WORLD_SIZE": str(world), "LOCAL_RANK": str(rank),
                       "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port)})
    sys.path.insert(0, ROOT)
    from kubernetes_acs_engine_autoscaler_b200 import distributed as D
    r, w, _ = D.init("gloo")
    assert (r, w) == (rank, world)
    c, res = shard_tick(rank)
    delta, counters = D.fleet_scale_up(res, c["pool_actual"])
    D.barrier()
    np.save(os.path.join(out_dir, "r%d.npy" % rank), np.concatenate([delta, counters]))
    D.shutdown()


def test_two_rank_fleet_totals(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    expect = None
    for rank in range(2):
        c, res = shard_tick(rank)
        v = np.concatenate([np.asarray(res["new_size"], np.int64) - c["pool_actual"].astype(np.int64),
                            np.array([res["n_to_schedule"], res["n_pending"], res["num_unaccounted"], res["decisions"]],
                                     np.int64)])
        expect = v if expect is None else expect + v
    for rank in range(2):
        got = np.load(os.path.join(str(tmp_path), "r%d.npy" % rank))
        np.testing.assert_array_equal(got, expect)
