#!/usr/bin/env python3
"""TEST INFRASTRUCTURE -- records mid-size ticks of the UNMODIFIED reference into tests/golden/mid_ticks.json.gz.

The fixtures of make_golden.py stop at 24 nodes x 280 pods; these cross what the CUDA pipeline is built from:
a 256-node stage, several 256-pod tiles, a second pool, D = 8 / T = 8 (generated CAPACITY_DATA), the max_size
raise path and an all-fit tick.  Each case names the arguments of synthetic.make_cluster that regenerate its
snapshot (plus a digest of the generated arrays, so generator drift is caught) and stores what the reference
computed: which node took each pod to schedule (-1 = pending), node.used_capacity as float64 bits, the
scale_pools argument or the exception, and every node's get_node_state.  ~25 s of reference time.

    python oracle/make_golden_mid.py        (needs /root/reference or oracle/_ref)
"""
import base64
import gzip
import hashlib
import json
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import ref_bench  # noqa: E402
from kubernetes_acs_engine_autoscaler_b200 import synthetic as syn  # noqa: E402

CASES = [
    {"name": "1k_x_100_d4", "P": 1000, "N": 100, "D": 4, "T": 1, "seed": 101},
    {"name": "2k_x_300_d4_two_pools", "P": 2000, "N": 300, "D": 4, "T": 2, "seed": 102, "over_provision": 2},
    {"name": "4k_x_400_d4", "P": 4000, "N": 400, "D": 4, "T": 1, "seed": 103},
    {"name": "1500_x_300_d8_t8", "P": 1500, "N": 300, "D": 8, "T": 8, "seed": 104},
    {"name": "raise_path_max_size", "P": 800, "N": 60, "D": 4, "T": 2, "seed": 105, "max_size": 70},
    {"name": "all_fit", "P": 300, "N": 400, "D": 4, "T": 1, "seed": 106, "free_frac": 1.0, "run_per_node": 2},
    {"name": "gpu_heavy_d4_t8", "P": 1200, "N": 280, "D": 4, "T": 8, "seed": 107, "gpu_prob": 0.5},
]
STATE_CODES = {"instance-terminated": 0, "pod-pending": 1, "grace-period": 2, "spare-agent": 3, "idle-schedulable": 4,
               "idle-unschedulable": 5, "busy-unschedulable": 6, "busy": 7, "under-utilized-drainable": 8,
               "under-utilized-undrainable": 9}


def cluster_of(case):
    kw = {k: case[k] for k in ("free_frac", "run_per_node", "gpu_prob", "max_size", "over_provision") if k in case}
    return syn.make_cluster(case["P"], case["N"], case["D"], case["T"], seed=case["seed"], **kw)


def digest(c):
    h = hashlib.sha256()
    for k in ("req", "cap_type", "node_type", "row_ptr", "run_idx", "req_run", "flags_run", "node_flags", "node_age"):
        h.update(np.ascontiguousarray(c[k]).tobytes())
    return h.hexdigest()


def pack(a):
    return base64.b64encode(zlib.compress(np.ascontiguousarray(a).tobytes(), 9)).decode("ascii")


def unpack(s, dtype):
    return np.frombuffer(zlib.decompress(base64.b64decode(s)), dtype=dtype)


def main():
    import oracle
    oracle.build()
    out = []
    for case in CASES:
        c = cluster_of(case)
        tick = ref_bench.ReferenceTick(c, max_size=case.get("max_size"))
        res = tick.run(record=True)
        dense = ref_bench.dense_answer(oracle, c)
        mism = ref_bench.compare(c, res, dense)
        assert not mism, (case["name"], mism)  # the C oracle agrees with the reference on this case, live
        to_schedule = np.array([int(u.split("-")[1]) for u in res["to_schedule"]], dtype=np.int32)
        rec = dict(case)
        rec.update({
            "digest": digest(c),
            "to_schedule": pack(to_schedule),
            "placed": pack(np.array(res["placed"], dtype=np.int32)),
            "used_bits": pack(res["used"].view(np.uint64)),
            "scale_pools": res["scale_calls"],
            "exception": res["exception"],
            "states": pack(np.array([STATE_CODES[s] for s in res["states"]], dtype=np.uint8)),
            "reference_seconds": round(sum(res["seconds"].values()), 3),
        })
        out.append(rec)
        print(case["name"], "to schedule", len(to_schedule), "placed", int((np.array(res["placed"]) >= 0).sum()),
              "scale", res["scale_calls"], "exc", res["exception"], "%.1fs" % rec["reference_seconds"])
    path = os.path.join(os.path.dirname(HERE), "tests", "golden", "mid_ticks.json.gz")
    with gzip.GzipFile(path, "wb", mtime=0) as f:
        f.write(json.dumps({"source": "unmodified reference via oracle/ref_bench.py", "cases": out},
                           sort_keys=True).encode())
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
