"""Device engine: thin Python face of libacsfit.so on dense float64 snapshots.

PyTorch is only the buffer container here (device memory, streams); every computation is a
hand-written sm_100a kernel behind the C ABI of include/acsfit.h.  There is no CPU
implementation: constructing an Engine without a CUDA device raises.
"""
import ctypes

import numpy as np
import torch

from . import _native
from ._native import AcsfitError

PIPELINE_DIMS = (2, 4, 8, 16)


def padded_dims(D):
    """columns the first-fit kernels are instantiated for; zero columns are neutral because an
    absent resource key is 0.0 on both sides of every test (kube.py:204-206, :210-212)."""
    for d in PIPELINE_DIMS:
        if D <= d:
            return d
    raise ValueError("more than %d resource dimensions are not supported" % PIPELINE_DIMS[-1])


def _ptr(t):
    if t is None:
        return None
    if isinstance(t, torch.Tensor):
        assert t.is_contiguous()
        return ctypes.c_void_p(t.data_ptr())
    assert isinstance(t, np.ndarray) and t.flags.c_contiguous
    return ctypes.c_void_p(t.ctypes.data)


def _np(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


class Engine(object):
    """one acsfit context on one GPU (one per process / per rank)."""

    def __init__(self, device=None, min_stages=0, watchdog_ms=0):
        self._lib = _native.load()
        if not torch.cuda.is_available():
            raise RuntimeError("acsfit needs a CUDA device (sm_100a); there is no CPU fallback")
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device("cuda", int(device))
        self._ctx = ctypes.c_void_p()
        st = self._lib.acsfit_ctx_create(int(device), ctypes.byref(self._ctx))
        if st != 0:
            raise AcsfitError(st, "acsfit_ctx_create failed on device %s" % device)
        if min_stages or watchdog_ms:
            self._check(self._lib.acsfit_ctx_configure(self._ctx, int(min_stages), int(watchdog_ms)))

    def close(self):
        if self._ctx:
            self._lib.acsfit_ctx_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _check(self, status):
        if status != 0:
            raise AcsfitError(status, self._lib.acsfit_last_error(self._ctx).decode())

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def dev(self, a, dtype):
        """host array -> contiguous device tensor (H2D through pinned memory when large)."""
        if isinstance(a, torch.Tensor):
            return a.to(device=self.device, dtype=dtype).contiguous()
        t = torch.from_numpy(np.ascontiguousarray(a))
        return t.to(device=self.device, dtype=dtype, non_blocking=False).contiguous()

    def empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype, device=self.device)

    def set_timing(self, enabled):
        self._check(self._lib.acsfit_ctx_set_timing(self._ctx, 1 if enabled else 0))

    def set_knob(self, name, value):
        """developer knobs of include/acsfit.h ("ranks", "prune", "overlap", "min_stages", "cluster_blocks")."""
        self._check(self._lib.acsfit_ctx_set_knob(self._ctx, name.encode(), int(value)))

    def configure(self, min_stages=0, watchdog_ms=0):
        self._check(self._lib.acsfit_ctx_configure(self._ctx, int(min_stages), int(watchdog_ms)))

    # ------------------------------------------------------------------ cluster mode (one cluster, all GPUs of a box)
    def cluster_connect(self, max_pods, max_nodes, max_dims=16):
        """join the ranks of the default torch.distributed group into ONE cluster (include/acsfit.h, "Cluster
        mode"): every rank exports its exchange region (CUDA IPC), the handles are all-gathered here and opened
        by the library.  Afterwards first_fit_nodes / fulfill_pending / scale_up / scale_up_host must be called by
        all ranks with identical (replicated) inputs; each returns the complete result."""
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        handle = (ctypes.c_ubyte * 64)()
        self._check(self._lib.acsfit_cluster_init(self._ctx, rank, world, int(max_pods), int(max_nodes),
                                                  int(max_dims), ctypes.cast(handle, ctypes.c_void_p)))
        where = self.device if dist.get_backend() == "nccl" else torch.device("cpu")
        mine = torch.tensor(list(bytes(handle)), dtype=torch.uint8, device=where)
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        blob = b"".join(bytes(t.cpu().numpy().tobytes()) for t in gathered)
        buf = ctypes.create_string_buffer(blob, len(blob))
        self._check(self._lib.acsfit_cluster_connect(self._ctx, ctypes.cast(buf, ctypes.c_void_p)))
        dist.barrier()  # every rank has zeroed and mapped every region before the first in-stream barrier
        self.cluster_world = world
        return world

    def cluster_barrier(self):
        self._check(self._lib.acsfit_cluster_barrier(self._ctx, self._stream()))

    def cluster_last_plan(self):
        v = [ctypes.c_int() for _ in range(4)]
        self._check(self._lib.acsfit_cluster_last_plan(self._ctx, *[ctypes.cast(ctypes.byref(x), ctypes.c_void_p) for x in v]))
        return {"stage_nodes": v[0].value, "stages": v[1].value, "pod_blocks": v[2].value, "resident": v[3].value}

    @property
    def launch_count(self):
        return int(self._lib.acsfit_launch_count(self._ctx))

    def pipeline_stats(self):
        ms = ctypes.c_double()
        dec = ctypes.c_uint64()
        stages = ctypes.c_int()
        tiles = ctypes.c_int()
        self._check(self._lib.acsfit_last_pipeline_stats(self._ctx, ctypes.byref(ms), ctypes.byref(dec),
                                                         ctypes.byref(stages), ctypes.byref(tiles)))
        return {"ms": ms.value, "decisions": dec.value, "stages": stages.value, "tiles": tiles.value}

    def debug_profile(self, enabled=True):
        """developer probe: per-stage clock64 totals [stages, 8] of the last pipeline launch
        (wait, load, scan, resolve, publish, refresh, hits, tiles)."""
        out = np.zeros((4096, 8), dtype=np.uint64)
        n = ctypes.c_int(0)
        self._check(self._lib.acsfit_debug_profile(self._ctx, 1 if enabled else 0, _ptr(out), 4096,
                                                   ctypes.cast(ctypes.byref(n), ctypes.c_void_p)))
        return out[:n.value]

    def debug_trace(self, stage, tiles=8192):
        """developer probe: select the stage to trace / fetch its per-tile trace [tiles, 8]."""
        out = np.zeros((tiles, 8), dtype=np.uint64)
        self._check(self._lib.acsfit_debug_trace(self._ctx, int(stage), _ptr(out), tiles))
        return out

    # ------------------------------------------------------------------ device entry points
    def feasible_mask(self, req, unit):
        """capacity.is_possible over all pools (capacity.py:24-32). req [P,D], unit [T,D] device."""
        P, D = req.shape
        T = unit.shape[0]
        mask = self.empty((P,), torch.uint8)
        evals = torch.zeros(1, dtype=torch.int64, device=self.device)
        self._check(self._lib.acsfit_feasible_mask(self._ctx, _ptr(req), P, D, _ptr(unit), T, _ptr(mask),
                                                   _ptr(evals), self._stream()))
        return mask, evals

    def occupancy(self, row_ptr, run_idx, req_run, used):
        """cluster.py:165-168: ordered used += req per node; `used` [N,D] device, in place."""
        N, D = used.shape
        self._check(self._lib.acsfit_occupancy(self._ctx, _ptr(row_ptr), _ptr(run_idx), _ptr(req_run), N, D,
                                               _ptr(used), self._stream()))
        return used

    def first_fit_nodes(self, req, pod_idx, cap_type, node_type, used):
        """Cluster.get_pending_pods (cluster.py:184-204). Returns (placed int32[P] device,
        decisions int64[1] device); `used` is updated in place."""
        rows, D = req.shape
        P = rows if pod_idx is None else pod_idx.shape[0]
        N = used.shape[0]
        placed = self.empty((P,), torch.int32)
        decisions = torch.zeros(1, dtype=torch.int64, device=self.device)
        self._check(self._lib.acsfit_first_fit_nodes(self._ctx, _ptr(req), rows, _ptr(pod_idx), P, D, _ptr(cap_type),
                                                     _ptr(node_type), _ptr(used), N, _ptr(placed), _ptr(decisions),
                                                     self._stream()))
        return placed, decisions

    def fulfill_pending(self, req, num_listed, unit, pool_actual, pool_max, pool_ignored, over_provision):
        """Scaler.fulfill_pending up to scaler.py:177. req [Pp,D] device; pool arrays host (visiting order)."""
        Pp, D = req.shape
        unit = _np(unit, np.float64).reshape(-1, D)
        T = unit.shape[0]
        pool_actual = _np(pool_actual, np.int32)
        pool_max = _np(pool_max, np.int32)
        pool_ignored = _np(pool_ignored, np.uint8)
        new_size = np.zeros(T, dtype=np.int64)
        units_needed = np.zeros(T, dtype=np.int64)
        bins_opened = np.zeros(T, dtype=np.int64)
        acc_pool = self.empty((Pp,), torch.int32)
        bin_of = self.empty((Pp,), torch.int32)
        unacc = ctypes.c_int64(0)
        evals = ctypes.c_uint64(0)
        self._check(self._lib.acsfit_fulfill_pending(
            self._ctx, _ptr(req), Pp, int(num_listed), D, _ptr(unit), _ptr(pool_actual), _ptr(pool_max),
            _ptr(pool_ignored), T, int(over_provision), _ptr(new_size), _ptr(units_needed), _ptr(bins_opened),
            _ptr(acc_pool), _ptr(bin_of), ctypes.cast(ctypes.byref(unacc), ctypes.c_void_p),
            ctypes.cast(ctypes.byref(evals), ctypes.c_void_p), self._stream()))
        return {"new_size": new_size, "units_needed": units_needed, "bins_opened": bins_opened,
                "acc_pool": acc_pool, "bin_of": bin_of, "num_unaccounted": int(unacc.value),
                "evals": int(evals.value)}

    def node_states(self, row_ptr, run_idx, req_run, flags_run, cap_type, node_type, node_flags, node_age,
                    any_pending, idle_thresholds):
        """Scaler.get_node_state for all nodes x S thresholds (scaler.py:61-114). returns uint8 [S,N] device."""
        N = node_type.shape[0]
        D = cap_type.shape[1]
        thr = _np(idle_thresholds, np.int64)
        S = thr.shape[0]
        out = self.empty((S, N), torch.uint8)
        self._check(self._lib.acsfit_node_states(self._ctx, _ptr(row_ptr), _ptr(run_idx), _ptr(req_run),
                                                 _ptr(flags_run), _ptr(cap_type), _ptr(node_type), _ptr(node_flags),
                                                 _ptr(node_age), N, D, 1 if any_pending else 0, _ptr(thr), S,
                                                 _ptr(out), self._stream()))
        return out

    def maintain_actions(self, state, node_pool, budget0, pool_scalable, dry_run):
        """engine_scaler.py:133-182 decisions. `state` uint8[N] device is rewritten; returns action uint8[N]."""
        N = state.shape[0]
        budget0 = _np(budget0, np.int64)
        pool_scalable = _np(pool_scalable, np.uint8)
        T = budget0.shape[0]
        action = self.empty((N,), torch.uint8)
        self._check(self._lib.acsfit_maintain_actions(self._ctx, _ptr(state), _ptr(node_pool), N, _ptr(budget0),
                                                      _ptr(pool_scalable), T, 1 if dry_run else 0, _ptr(action),
                                                      self._stream()))
        return state, action

    def scale_up(self, req, unit_all, unit_ordered, pool_actual, pool_max, pool_ignored, over_provision,
                 cap_type, node_type, used):
        """fused get_pods_to_schedule + get_pending_pods + fulfill_pending on device buffers."""
        P, D = req.shape
        N = used.shape[0]
        unit_all = _np(unit_all, np.float64).reshape(-1, D)
        unit_ordered = _np(unit_ordered, np.float64).reshape(-1, D)
        T = unit_all.shape[0]
        pool_actual = _np(pool_actual, np.int32)
        pool_max = _np(pool_max, np.int32)
        pool_ignored = _np(pool_ignored, np.uint8)
        feasible = self.empty((P,), torch.uint8)
        placed = self.empty((P,), torch.int32)
        acc_pool = self.empty((P,), torch.int32)
        new_size = np.zeros(T, dtype=np.int64)
        units_needed = np.zeros(T, dtype=np.int64)
        bins_opened = np.zeros(T, dtype=np.int64)
        counters = np.zeros(4, dtype=np.uint64)
        self._check(self._lib.acsfit_scale_up(
            self._ctx, _ptr(req), P, D, _ptr(unit_all), _ptr(unit_ordered), _ptr(pool_actual), _ptr(pool_max),
            _ptr(pool_ignored), T, int(over_provision), _ptr(cap_type), _ptr(node_type), _ptr(used), N,
            _ptr(feasible), _ptr(placed), _ptr(new_size), _ptr(units_needed), _ptr(bins_opened), _ptr(acc_pool),
            _ptr(counters), self._stream()))
        return {"feasible": feasible, "placed": placed, "acc_pool": acc_pool, "new_size": new_size,
                "units_needed": units_needed, "bins_opened": bins_opened,
                "n_to_schedule": int(counters[0]), "n_pending": int(counters[1]),
                "num_unaccounted": int(counters[2]), "decisions": int(counters[3])}

    # ------------------------------------------------------------------ host-buffer entry points
    def scale_up_host(self, req, unit_all, unit_ordered, pool_actual, pool_max, pool_ignored, over_provision,
                      cap_type, node_type, used, out=None):
        """the plugin call: all buffers are HOST numpy arrays (pinned for speed); `used` is updated in place."""
        req = _np(req, np.float64)
        P, D = req.shape
        cap_type = _np(cap_type, np.float64).reshape(-1, D)
        K = cap_type.shape[0]
        node_type = _np(node_type, np.int32)
        assert isinstance(used, np.ndarray) and used.dtype == np.float64 and used.flags.c_contiguous
        N = used.shape[0]
        unit_all = _np(unit_all, np.float64).reshape(-1, D)
        unit_ordered = _np(unit_ordered, np.float64).reshape(-1, D)
        T = unit_all.shape[0]
        pool_actual = _np(pool_actual, np.int32)
        pool_max = _np(pool_max, np.int32)
        pool_ignored = _np(pool_ignored, np.uint8)
        if out is None:
            out = {"feasible": np.empty(P, dtype=np.uint8), "placed": np.empty(P, dtype=np.int32),
                   "acc_pool": np.empty(P, dtype=np.int32)}
        new_size = np.zeros(T, dtype=np.int64)
        units_needed = np.zeros(T, dtype=np.int64)
        bins_opened = np.zeros(T, dtype=np.int64)
        counters = np.zeros(4, dtype=np.uint64)
        self._check(self._lib.acsfit_scale_up_host(
            self._ctx, _ptr(req), P, D, _ptr(unit_all), _ptr(unit_ordered), _ptr(pool_actual), _ptr(pool_max),
            _ptr(pool_ignored), T, int(over_provision), _ptr(cap_type), K, _ptr(node_type), _ptr(used), N,
            _ptr(out["feasible"]), _ptr(out["placed"]), _ptr(new_size), _ptr(units_needed), _ptr(bins_opened),
            _ptr(out["acc_pool"]), _ptr(counters)))
        res = dict(out)
        res.update({"new_size": new_size, "units_needed": units_needed, "bins_opened": bins_opened,
                    "n_to_schedule": int(counters[0]), "n_pending": int(counters[1]),
                    "num_unaccounted": int(counters[2]), "decisions": int(counters[3])})
        return res

    def occupancy_host(self, row_ptr, run_idx, req_run, used):
        """cluster.py:165-168 with HOST arrays; `used` [N, D] float64 is updated in place."""
        assert isinstance(used, np.ndarray) and used.dtype == np.float64 and used.flags.c_contiguous
        N, D = used.shape
        req_run = _np(req_run, np.float64).reshape(-1, D)
        self._check(self._lib.acsfit_occupancy_host(self._ctx, _ptr(_np(row_ptr, np.int64)),
                                                    _ptr(None if run_idx is None else _np(run_idx, np.int32)),
                                                    _ptr(req_run), req_run.shape[0], N, D, _ptr(used)))
        return used

    def maintain_host(self, row_ptr, run_idx, req_run, flags_run, cap_type, node_type, node_flags, node_age,
                      node_pool, any_pending, idle_threshold, budget0, pool_scalable, dry_run):
        cap_type = _np(cap_type, np.float64)
        K, D = cap_type.shape
        req_run = _np(req_run, np.float64).reshape(-1, D)
        R = req_run.shape[0]
        row_ptr = _np(row_ptr, np.int64)
        run_idx = None if run_idx is None else _np(run_idx, np.int32)  # None: contiguous table
        flags_run = _np(flags_run, np.uint8)
        node_type = _np(node_type, np.int32)
        node_flags = _np(node_flags, np.uint8)
        node_age = _np(node_age, np.int64)
        node_pool = _np(node_pool, np.int32)
        budget0 = _np(budget0, np.int64)
        pool_scalable = _np(pool_scalable, np.uint8)
        N = node_type.shape[0]
        T = budget0.shape[0]
        state = np.empty(N, dtype=np.uint8)
        action = np.empty(N, dtype=np.uint8)
        self._check(self._lib.acsfit_maintain_host(
            self._ctx, _ptr(row_ptr), _ptr(run_idx), _ptr(req_run), _ptr(flags_run), R, _ptr(cap_type), K,
            _ptr(node_type), _ptr(node_flags), _ptr(node_age), _ptr(node_pool), N, D, 1 if any_pending else 0,
            int(idle_threshold), _ptr(budget0), _ptr(pool_scalable), T, 1 if dry_run else 0, _ptr(state),
            _ptr(action)))
        return state, action
