"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink; gloo on CPU in tests).

The tick is a sequential recurrence over the pod list, so ONE cluster does not shard on the pod axis
(DESIGN.md section 5).  What scales today is a fleet of independent cluster shards: every rank runs the
whole tick for its shard on its own GPU and a single all-reduce(sum) of the int64 per-pool scale-up
counts (plus a few counters) yields the fleet totals.  No data-path collective.
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


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def shutdown():
    if dist.is_initialized():
        dist.destroy_process_group()
