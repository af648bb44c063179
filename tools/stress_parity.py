#!/usr/bin/env python3
"""developer stress run (GPU): many random cluster shapes through the fused tick, each compared bit for bit with
the plain-C oracle (placements, accounted pools, per-pool integers, decision count, `used` bit patterns).
`ACSFIT_PRUNE=1` / `ACSFIT_OVERLAP=0` in the environment exercise the alternative schedules.

  python tools/stress_parity.py --cases 300 --seed 1
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402  (test infrastructure: the checker)
from kubernetes_acs_engine_autoscaler_b200 import synthetic as syn  # noqa: E402
from kubernetes_acs_engine_autoscaler_b200.engine import Engine  # noqa: E402
from test_gpu_parity import oracle_scale_up  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=200)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    rng = np.random.default_rng(args.seed)
    eng = Engine(0)
    bad = 0
    for case in range(args.cases):
        D = int(rng.choice([4, 4, 8, 8]))
        T = int(rng.integers(1, 9)) if D == 8 else int(rng.integers(1, 4))
        P = int(rng.choice([1, 17, 255, 256, 257, 1000, 4000, 12000, 30000]))
        N = int(rng.choice([1, 31, 32, 33, 100, 257, 1000, 3000]))
        free = float(rng.choice([0.0, 0.1, 0.3, 0.6, 1.0]))
        max_size = None if rng.random() < 0.7 else int(rng.integers(1, 200))
        over = int(rng.integers(0, 6))
        seed = int(rng.integers(0, 2 ** 31))
        c = syn.make_cluster(P, N, D, T, seed=seed, free_frac=free, max_size=max_size, over_provision=over)
        used0 = syn.initial_used(c)
        used_o = used0.copy()
        o = oracle_scale_up(oracle, c, used_o)
        used_h = used0.copy()
        h = eng.scale_up_host(c["req"], c["unit_all"], c["unit_ordered"], c["pool_actual"], c["pool_max"],
                              c["pool_ignored"], c["over_provision"], c["cap_type"], c["node_type"], used_h)
        ok = True
        for k in ("feasible", "placed", "acc_pool", "new_size", "units_needed", "bins_opened"):
            ok = ok and np.array_equal(np.asarray(h[k]), np.asarray(o[k]))
        for k in ("n_to_schedule", "n_pending", "num_unaccounted", "decisions"):
            ok = ok and h[k] == o[k]
        ok = ok and used_h.tobytes() == used_o.tobytes()
        if not ok:
            bad += 1
            print("MISMATCH case %d: P=%d N=%d D=%d T=%d free=%.1f max=%s over=%d seed=%d" % (case, P, N, D, T, free, max_size, over, seed),
                  flush=True)
    print("%d cases, %d mismatches" % (args.cases, bad))
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
