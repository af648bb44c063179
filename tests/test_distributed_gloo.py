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


# ---- one cluster, nodes range-sharded over the ranks, pod blocks pipelined through them (exact) ----
def cluster_case():
    sys.path.insert(0, ROOT)
    from kubernetes_acs_engine_autoscaler_b200 import synthetic as syn
    c = syn.make_cluster(4000, 301, 4, 2, seed=77)
    return c, syn.initial_used(c)


def cluster_worker(rank, world, port, out_dir, n_blocks):
    os.environ.update({"RANK": str(rank), "WORLD_SIZE": str(world), "LOCAL_RANK": str(rank),
                       "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port)})
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    from kubernetes_acs_engine_autoscaler_b200 import distributed as D
    from oracle_engine import OracleEngine
    D.init("gloo")
    c, used0 = cluster_case()
    eng = OracleEngine()
    N = c["N"]
    lo, hi = (N * rank) // world, (N * (rank + 1)) // world
    req = eng.dev(c["req"], torch.float64)
    mask, _ = eng.feasible_mask(req, eng.dev(c["unit_all"], torch.float64))
    feas = torch.nonzero(mask).flatten().to(torch.int32)
    used = eng.dev(used0[lo:hi], torch.float64)
    placed, dec = D.cluster_first_fit(eng, req, feas, eng.dev(c["cap_type"], torch.float64),
                                      eng.dev(c["node_type"][lo:hi], torch.int32), used, lo, n_blocks=n_blocks)
    np.savez(os.path.join(out_dir, "c%d.npz" % rank), placed=placed.numpy(), dec=dec.numpy(), used=used.numpy(), lo=lo, hi=hi)
    D.barrier()
    D.shutdown()


def test_three_rank_cluster_pipeline_with_tiny_blocks(tmp_path):
    """uneven node ranges (301 nodes over 3 ranks), more blocks than some ranks ever see alive pods for"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    c, used0 = cluster_case()
    feas_mask, _ = oracle.feasible_mask(c["req"], c["unit_all"])
    rows = c["req"][np.nonzero(feas_mask)[0]]
    used_ref = used0.copy()
    placed_ref, calls_ref = oracle.first_fit_nodes(rows, c["cap_type"], c["node_type"], used_ref)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = os.path.join(str(tmp_path), "w3")
    os.makedirs(out)
    mp.spawn(cluster_worker, args=(3, port, out, 64), nprocs=3, join=True)
    used_got = np.zeros_like(used0)
    for rank in range(3):
        z = np.load(os.path.join(out, "c%d.npz" % rank))
        np.testing.assert_array_equal(z["placed"], placed_ref)
        assert int(z["dec"][0]) == calls_ref
        used_got[int(z["lo"]):int(z["hi"])] = z["used"]
    assert used_got.tobytes() == used_ref.tobytes()


def test_two_rank_cluster_pipeline_is_exact(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    c, used0 = cluster_case()
    feas_mask, _ = oracle.feasible_mask(c["req"], c["unit_all"])
    rows = c["req"][np.nonzero(feas_mask)[0]]
    used_ref = used0.copy()
    placed_ref, calls_ref = oracle.first_fit_nodes(rows, c["cap_type"], c["node_type"], used_ref)
    assert (placed_ref >= 0).sum() > 50 and (placed_ref < 0).sum() > 50   # both outcomes are exercised
    for n_blocks in (1, 7):
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        out = os.path.join(str(tmp_path), "b%d" % n_blocks)
        os.makedirs(out)
        mp.spawn(cluster_worker, args=(2, port, out, n_blocks), nprocs=2, join=True)
        used_got = np.zeros_like(used0)
        for rank in range(2):
            z = np.load(os.path.join(out, "c%d.npz" % rank))
            np.testing.assert_array_equal(z["placed"], placed_ref)
            assert int(z["dec"][0]) == calls_ref
            used_got[int(z["lo"]):int(z["hi"])] = z["used"]
        assert used_got.tobytes() == used_ref.tobytes()      # bit-exact node state
