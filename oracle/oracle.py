"""TEST INFRASTRUCTURE -- ctypes wrapper of oracle/libacsfit_oracle.so (acsfit_oracle.c).

The plain-C, single-threaded restatement of the reference's tick decision path.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module; the product never does.  Build with `make -C oracle`.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libacsfit_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "acsfit_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "libacsfit_oracle.so"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        for name in ("oracle_feasible_mask", "oracle_first_fit_nodes", "oracle_fulfill_pending",
                     "oracle_count_decisions"):
            getattr(_lib, name).restype = ctypes.c_uint64
        for name in ("oracle_occupancy", "oracle_node_states", "oracle_maintain_actions"):
            getattr(_lib, name).restype = None
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


def feasible_mask(req, unit):
    req = _f64(req)
    unit = _f64(unit)
    P, D = req.shape
    T = unit.shape[0]
    mask = np.zeros(P, dtype=np.uint8)
    evals = lib().oracle_feasible_mask(_p(req), ctypes.c_int64(P), ctypes.c_int(D), _p(unit), ctypes.c_int(T), _p(mask))
    return mask, int(evals)


def occupancy(row_ptr, run_idx, req_run, used):
    """in-place on `used` (N x D float64, C-contiguous)."""
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int64)
    run_idx = np.ascontiguousarray(run_idx, dtype=np.int32)
    req_run = _f64(req_run)
    assert used.dtype == np.float64 and used.flags.c_contiguous
    N, D = used.shape
    lib().oracle_occupancy(_p(row_ptr), _p(run_idx), _p(req_run), ctypes.c_int64(N), ctypes.c_int(D), _p(used))
    return used


def first_fit_nodes(req, cap_type, node_type, used):
    """sequential first fit; mutates `used` in place. returns (placed int32[P], can_fit calls)."""
    req = _f64(req)
    cap_type = _f64(cap_type)
    node_type = np.ascontiguousarray(node_type, dtype=np.int32)
    assert used.dtype == np.float64 and used.flags.c_contiguous
    P, D = req.shape
    N = used.shape[0]
    placed = np.full(P, -1, dtype=np.int32)
    calls = lib().oracle_first_fit_nodes(_p(req), ctypes.c_int64(P), ctypes.c_int(D), _p(cap_type), _p(node_type),
                                         _p(used), ctypes.c_int64(N), _p(placed))
    return placed, int(calls)


def fulfill_pending(req, num_listed, unit, pool_actual, pool_max, pool_ignored, over_provision):
    req = _f64(req)
    unit = _f64(unit)
    Pp, D = req.shape
    T = unit.shape[0]
    pool_actual = np.ascontiguousarray(pool_actual, dtype=np.int32)
    pool_max = np.ascontiguousarray(pool_max, dtype=np.int32)
    pool_ignored = np.ascontiguousarray(pool_ignored, dtype=np.uint8)
    new_size = np.zeros(T, dtype=np.int64)
    units_needed = np.zeros(T, dtype=np.int64)
    bins_opened = np.zeros(T, dtype=np.int64)
    acc_pool = np.full(Pp, -1, dtype=np.int32)
    bin_of = np.full(Pp, -1, dtype=np.int32)
    unacc = ctypes.c_int64(0)
    evals = lib().oracle_fulfill_pending(
        _p(req), ctypes.c_int64(Pp), ctypes.c_int64(num_listed), ctypes.c_int(D), _p(unit), _p(pool_actual),
        _p(pool_max), _p(pool_ignored), ctypes.c_int(T), ctypes.c_int64(over_provision), _p(new_size),
        _p(units_needed), _p(bins_opened), _p(acc_pool), _p(bin_of), ctypes.byref(unacc))
    return {"new_size": new_size, "units_needed": units_needed, "bins_opened": bins_opened,
            "acc_pool": acc_pool, "bin_of": bin_of, "num_unaccounted": int(unacc.value), "evals": int(evals)}


def node_states(row_ptr, run_idx, req_run, flags_run, cap_type, node_type, node_flags, node_age, any_pending,
                idle_thresholds):
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int64)
    run_idx = np.ascontiguousarray(run_idx, dtype=np.int32)
    req_run = _f64(req_run)
    flags_run = np.ascontiguousarray(flags_run, dtype=np.uint8)
    cap_type = _f64(cap_type)
    node_type = np.ascontiguousarray(node_type, dtype=np.int32)
    node_flags = np.ascontiguousarray(node_flags, dtype=np.uint8)
    node_age = np.ascontiguousarray(node_age, dtype=np.int64)
    thr = np.ascontiguousarray(idle_thresholds, dtype=np.int64)
    N = node_type.shape[0]
    D = cap_type.shape[1]
    S = thr.shape[0]
    out = np.zeros((S, N), dtype=np.uint8)
    lib().oracle_node_states(_p(row_ptr), _p(run_idx), _p(req_run), _p(flags_run), _p(cap_type), _p(node_type),
                             _p(node_flags), _p(node_age), ctypes.c_int64(N), ctypes.c_int(D),
                             ctypes.c_int(1 if any_pending else 0), _p(thr), ctypes.c_int(S), _p(out))
    return out


def maintain_actions(state, node_pool, budget0, pool_scalable, dry_run):
    state = np.array(state, dtype=np.uint8, copy=True)
    node_pool = np.ascontiguousarray(node_pool, dtype=np.int32)
    budget0 = np.ascontiguousarray(budget0, dtype=np.int64)
    pool_scalable = np.ascontiguousarray(pool_scalable, dtype=np.uint8)
    N = state.shape[0]
    T = budget0.shape[0]
    action = np.zeros(N, dtype=np.uint8)
    lib().oracle_maintain_actions(_p(state), _p(node_pool), ctypes.c_int64(N), _p(budget0), _p(pool_scalable),
                                  ctypes.c_int(T), ctypes.c_int(1 if dry_run else 0), _p(action))
    return state, action


def count_decisions(placed, N):
    placed = np.ascontiguousarray(placed, dtype=np.int32)
    return int(lib().oracle_count_decisions(_p(placed), ctypes.c_int64(placed.shape[0]), ctypes.c_int64(N)))
