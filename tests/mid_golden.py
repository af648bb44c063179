"""Shared loader / checker of tests/golden/mid_ticks.json.gz (recorded from the unmodified reference by
oracle/make_golden_mid.py): regenerates each case's snapshot from its make_cluster arguments, verifies the digest
of the generated arrays, and compares a dense tick result with what the reference computed."""
import base64
import gzip
import hashlib
import json
import os
import zlib

import numpy as np

from kubernetes_acs_engine_autoscaler_b200 import synthetic as syn

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mid_ticks.json.gz")


def _unpack(s, dtype):
    return np.frombuffer(zlib.decompress(base64.b64decode(s)), dtype=dtype).copy()


def load_cases():
    with gzip.open(PATH, "rb") as f:
        return json.loads(f.read().decode())["cases"]


def case_ids():
    return [c["name"] for c in load_cases()]


def cluster_of(case):
    kw = {k: case[k] for k in ("free_frac", "run_per_node", "gpu_prob", "max_size", "over_provision") if k in case}
    c = syn.make_cluster(case["P"], case["N"], case["D"], case["T"], seed=case["seed"], **kw)
    h = hashlib.sha256()
    for k in ("req", "cap_type", "node_type", "row_ptr", "run_idx", "req_run", "flags_run", "node_flags", "node_age"):
        h.update(np.ascontiguousarray(c[k]).tobytes())
    assert h.hexdigest() == case["digest"], "synthetic generator drifted: regenerate the mid-size goldens"
    return c


def expected(case):
    return {"to_schedule": _unpack(case["to_schedule"], np.int32), "placed": _unpack(case["placed"], np.int32),
            "used_bits": _unpack(case["used_bits"], np.uint64).reshape(case["N"], case["D"]),
            "states": _unpack(case["states"], np.uint8), "scale_pools": case["scale_pools"],
            "exception": case["exception"]}


def check_tick(case, placed_all, used, new_size, num_unaccounted, states):
    """placed_all [P]: node, -1 pending, -2 infeasible; used [N, D] float64; new_size [T]; states uint8 [N]."""
    want = expected(case)
    placed_all = np.asarray(placed_all)
    feasible = np.nonzero(placed_all != -2)[0].astype(np.int32)
    np.testing.assert_array_equal(feasible, want["to_schedule"])
    np.testing.assert_array_equal(placed_all[feasible], want["placed"])
    np.testing.assert_array_equal(np.ascontiguousarray(used, dtype=np.float64).view(np.uint64), want["used_bits"])
    n_pending = int((want["placed"] < 0).sum())
    if want["exception"] is not None:
        assert num_unaccounted > 0 and want["scale_pools"] == []   # the reference raised before scale_pools
    elif n_pending:
        assert num_unaccounted == 0
        assert want["scale_pools"] == [{"pool%d" % t: int(new_size[t]) for t in range(case["T"])}]
    else:
        assert want["scale_pools"] == []                           # fulfill_pending is not called (cluster.py:214)
    np.testing.assert_array_equal(np.asarray(states, dtype=np.uint8), want["states"])
