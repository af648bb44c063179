"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink; gloo on CPU in tests).

The tick is a sequential recurrence over the pod list, so ONE cluster does not shard on the pod axis
(DESIGN.md section 5).  Two modes:

* fleet (`fleet_scale_up`): independent cluster shards, every rank runs the whole tick for its shard on its
  own GPU and a single all-reduce(sum) of the int64 per-pool scale-up counts (plus a few counters) yields
  the fleet totals.  No data-path collective.  This is what `bench.py --gpus N` measures.
* one cluster, exact (`cluster_first_fit`): the NODE list is range-sharded over the ranks and the pod list is
  cut into blocks; block b visits rank 0, 1, 2, ... in node order while rank r is already working on block
  b+1 (pipeline parallelism).  The only exchange is the block's "still unplaced" byte mask, sent point to
  point to the next rank.  Every rank sees its pods in list order and every pod sees the nodes in list order,
  so placements, `used` and the credited decision count are bit-identical to the single-GPU result.
"""
import os

import numpy as np
import torch
import torch.distributed as dist


def env_rank():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init(backend=None):
    """idempotent process-group setup from the torchrun environment."""
    rank, world, local_rank = env_rank()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend)
    return rank, world, local_rank


def allreduce_counts(values, device=None):
    """sum an int64 vector over all ranks (the per-pool integer counts of the north-star)."""
    v = np.ascontiguousarray(values, dtype=np.int64)
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return v.copy()
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.from_numpy(v).to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def fleet_scale_up(tick_result, pool_actual):
    """per-rank tick result (engine.scale_up*/oracle dict) -> fleet totals over all cluster shards:
    (per-pool scale-up counts, [pods to schedule, pending, unaccounted, decisions])."""
    delta = np.asarray(tick_result["new_size"], dtype=np.int64) - np.asarray(pool_actual, dtype=np.int64)
    counters = np.array([tick_result["n_to_schedule"], tick_result["n_pending"], tick_result["num_unaccounted"],
                         tick_result["decisions"]], dtype=np.int64)
    total = allreduce_counts(np.concatenate([delta, counters]))
    return total[:len(delta)], total[len(delta):]


def _block_bounds(n, n_blocks):
    n_blocks = max(1, min(int(n_blocks), max(1, n)))
    edges = [(n * b) // n_blocks for b in range(n_blocks + 1)]
    return [(edges[b], edges[b + 1]) for b in range(n_blocks)]


def cluster_first_fit(engine, req, pod_idx, cap_type, node_type_local, used_local, node_offset, n_blocks=None):
    """Cluster.get_pending_pods (reference cluster.py:184-204) for ONE cluster whose nodes are range-sharded
    over the ranks: this rank owns nodes [node_offset, node_offset + len(node_type_local)) with their
    `used_local` rows (updated in place); `req` [P, D] and the ordered pod list `pod_idx` (int32 rows of req)
    are replicated.  Returns (placed, decisions): placed[i] = GLOBAL node index of list entry i or -1
    (identical on every rank), decisions = the reference's can_fit evaluation count summed over the ranks.
    """
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    device = req.device
    M = int(pod_idx.shape[0])
    placed = torch.full((M,), -1, dtype=torch.int32, device=device)
    decisions = torch.zeros(1, dtype=torch.int64, device=device)
    if n_blocks is None:
        n_blocks = 4 * world if world > 1 else 1
    pending_sends = []
    for (s, e) in _block_bounds(M, n_blocks):
        if e == s:
            continue
        if rank > 0:
            alive_u8 = torch.empty(e - s, dtype=torch.uint8, device=device)
            dist.recv(alive_u8, src=rank - 1)
            alive = alive_u8.bool()
        else:
            alive = torch.ones(e - s, dtype=torch.bool, device=device)
        idx = pod_idx[s:e][alive].contiguous()
        still = alive.clone()
        if idx.numel():
            placed_local, dec = engine.first_fit_nodes(req, idx, cap_type, node_type_local, used_local)
            placed_local = placed_local.to(device=device, dtype=torch.int32)
            decisions += dec.to(device=device, dtype=torch.int64).reshape(-1)[:1]
            hit = placed_local >= 0
            block = placed[s:e]
            block[alive] = torch.where(hit, placed_local + int(node_offset), torch.full_like(placed_local, -1))
            still[alive] = ~hit
        if rank < world - 1:
            out = still.to(torch.uint8).contiguous()
            pending_sends.append((dist.isend(out, dst=rank + 1), out))  # keep the buffer alive until sent
    for work, _ in pending_sends:
        work.wait()
    if world > 1:
        dist.all_reduce(placed, op=dist.ReduceOp.MAX)   # exactly one rank holds a value >= 0 per placed pod
        dist.all_reduce(decisions, op=dist.ReduceOp.SUM)
    return placed, decisions


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def shutdown():
    if dist.is_initialized():
        dist.destroy_process_group()
