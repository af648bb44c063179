"""GPU parity tests proper: the CUDA path (through the C ABI) against the plain-C oracle on the
same seeded inputs.  Bit-exact: integer outputs equal, float64 state compared as uint64 patterns."""
import numpy as np
import pytest
import torch

from kubernetes_acs_engine_autoscaler_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def to_np(t):
    return t.detach().cpu().numpy()


def run_first_fit(engine, oracle_mod, req, cap_type, node_type, used0, pod_idx=None):
    used_o = used0.copy()
    rows = req if pod_idx is None else req[pod_idx]
    placed_o, calls_o = oracle_mod.first_fit_nodes(rows, cap_type, node_type, used_o)
    d_req = engine.dev(req, torch.float64)
    d_used = engine.dev(used0, torch.float64)
    d_idx = None if pod_idx is None else engine.dev(pod_idx, torch.int32)
    placed, dec = engine.first_fit_nodes(d_req, d_idx, engine.dev(cap_type, torch.float64),
                                         engine.dev(node_type, torch.int32), d_used)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(to_np(placed), placed_o)
    np.testing.assert_array_equal(bits(to_np(d_used)), bits(used_o))
    assert int(to_np(dec)[0]) == calls_o
    return placed_o


@pytest.mark.parametrize("P,N,D,T,seed,free", [
    (16, 4, 4, 1, 0, 0.25), (300, 37, 4, 1, 1, 0.3), (1000, 100, 4, 1, 2, 0.2), (5000, 700, 4, 2, 3, 0.25),
    (4097, 1025, 8, 8, 4, 0.25), (20000, 3000, 4, 1, 5, 0.15), (20000, 5000, 8, 8, 6, 0.2),
    (257, 1, 4, 1, 7, 1.0), (1, 300, 4, 1, 8, 0.5),
])
def test_first_fit_nodes_matches_oracle(engine, oracle_mod, P, N, D, T, seed, free):
    c = syn.make_cluster(P, N, D, T, seed=seed, free_frac=free)
    used0 = syn.initial_used(c)
    run_first_fit(engine, oracle_mod, c["req"], c["cap_type"], c["node_type"], used0)


def test_first_fit_extremes(engine, oracle_mod):
    c = syn.make_cluster(3000, 400, 4, 1, seed=11)
    # all-fit: empty nodes, every early pod lands on the first nodes (long commit chains)
    run_first_fit(engine, oracle_mod, c["req"], c["cap_type"], c["node_type"], np.zeros((400, 4)))
    # none-fit: every node exactly full
    full = np.repeat(c["cap_type"][c["node_type"]], 1, axis=0).copy()
    placed = run_first_fit(engine, oracle_mod, c["req"], c["cap_type"], c["node_type"], full)
    assert (placed < 0).all()
    # zero nodes / zero pods
    d_req = engine.dev(c["req"], torch.float64)
    placed, dec = engine.first_fit_nodes(d_req, None, engine.dev(c["cap_type"], torch.float64),
                                         engine.empty((0,), torch.int32), engine.empty((0, 4), torch.float64))
    assert (to_np(placed) == -1).all() and int(to_np(dec)[0]) == 0
    placed, dec = engine.first_fit_nodes(engine.empty((0, 4), torch.float64), None,
                                         engine.dev(c["cap_type"], torch.float64),
                                         engine.dev(c["node_type"], torch.int32),
                                         engine.dev(np.zeros((400, 4)), torch.float64))
    assert placed.numel() == 0


def test_first_fit_with_pod_list_and_dims(engine, oracle_mod):
    rng = np.random.default_rng(5)
    for D in (2, 4, 8, 16):
        P, N, K = 1500, 211, 3
        cap = rng.integers(1, 40, size=(K, D)).astype(np.float64)
        req = (rng.integers(0, 4000, size=(P, D)).astype(np.float64) * 1e-3)
        req[rng.random((P, D)) < 0.3] = 0.0
        node_type = rng.integers(0, K, size=N).astype(np.int32)
        used0 = cap[node_type] * rng.random((N, D)) * (rng.random((N, 1)) < 0.8)
        idx = np.sort(rng.choice(P, size=P // 2, replace=False)).astype(np.int32)
        run_first_fit(engine, oracle_mod, req, cap, node_type, used0, pod_idx=idx)
        run_first_fit(engine, oracle_mod, req, cap, node_type, used0)


def test_first_fit_rounding_sensitive_values(engine, oracle_mod):
    # quantities whose sums round: 0.1-ish millicore values against capacities hit exactly / by one ulp
    rng = np.random.default_rng(9)
    P, N, D = 4000, 64, 4
    req = np.zeros((P, D))
    req[:, 1] = rng.integers(1, 30, size=P).astype(np.float64) * 1e-3 * 100
    req[:, 3] = 1.0
    cap = np.array([[0.0, 2.0, 7096762368.0, 110.0], [0.0, np.nextafter(2.0, 3.0), 1e9, 30.0]])
    node_type = (np.arange(N) % 2).astype(np.int32)
    run_first_fit(engine, oracle_mod, req, cap, node_type, np.zeros((N, D)))


def fulfill_both(engine, oracle_mod, req, num_listed, unit, actual, pmax, ignored, over):
    o = oracle_mod.fulfill_pending(req, num_listed, unit, actual, pmax, ignored, over)
    g = engine.fulfill_pending(engine.dev(req, torch.float64), num_listed, unit, actual, pmax, ignored, over)
    torch.cuda.synchronize()
    for k in ("new_size", "units_needed", "bins_opened"):
        np.testing.assert_array_equal(g[k], o[k], err_msg=k)
    np.testing.assert_array_equal(to_np(g["acc_pool"]), o["acc_pool"])
    np.testing.assert_array_equal(to_np(g["bin_of"]), o["bin_of"])
    assert g["num_unaccounted"] == o["num_unaccounted"]
    assert g["evals"] == o["evals"]
    return o


@pytest.mark.parametrize("P,D,T,seed,max_size,over,ignored", [
    (16, 4, 1, 0, None, 0, ()), (700, 4, 1, 1, None, 1, ()), (5000, 4, 2, 2, None, 0, ()),
    (5000, 8, 8, 3, None, 2, (1,)), (3000, 4, 3, 4, 100, 0, ()), (3000, 8, 8, 5, 40, 1, (0, 2)),
    (20000, 4, 1, 6, None, 0, ()), (1, 4, 1, 7, None, 0, ()), (9000, 8, 4, 8, 300, 0, ()),
])
def test_fulfill_pending_matches_oracle(engine, oracle_mod, P, D, T, seed, max_size, over, ignored):
    c = syn.make_cluster(P, 64, D, T, seed=seed, max_size=max_size, ignored=ignored)
    fulfill_both(engine, oracle_mod, c["req"], P, c["unit_ordered"], c["pool_actual"], c["pool_max"],
                 c["pool_ignored"], over)


def test_fulfill_pending_edge_cases(engine, oracle_mod):
    c = syn.make_cluster(500, 10, 4, 2, seed=3)
    # duplicates counted in num_listed but collapsed in req (scaler.py:119-120) -> never fully accounted
    base = fulfill_both(engine, oracle_mod, c["req"], 500, c["unit_ordered"], c["pool_actual"], c["pool_max"],
                        c["pool_ignored"], 0)
    o = fulfill_both(engine, oracle_mod, c["req"], 503, c["unit_ordered"], c["pool_actual"], c["pool_max"],
                     c["pool_ignored"], 0)
    assert o["num_unaccounted"] == base["num_unaccounted"] + 3
    # actual > max_size -> negative units_requested
    fulfill_both(engine, oracle_mod, c["req"], 500, c["unit_ordered"], np.array([7, 9], np.int32),
                 np.array([3, 100], np.int32), c["pool_ignored"], 1)
    # all pools ignored; zero pods
    fulfill_both(engine, oracle_mod, c["req"], 500, c["unit_ordered"], c["pool_actual"], c["pool_max"],
                 np.ones(2, np.uint8), 0)
    fulfill_both(engine, oracle_mod, np.zeros((0, 4)), 0, c["unit_ordered"], c["pool_actual"], c["pool_max"],
                 c["pool_ignored"], 2)
    # zero-request rows (every pod fits bin 0 forever)
    fulfill_both(engine, oracle_mod, np.zeros((300, 4)), 300, c["unit_ordered"], c["pool_actual"], c["pool_max"],
                 c["pool_ignored"], 0)


def test_feasible_mask_and_occupancy(engine, oracle_mod):
    for (P, N, D, T, seed) in [(1000, 50, 4, 1, 0), (7000, 900, 8, 8, 1), (33, 5, 4, 3, 2)]:
        c = syn.make_cluster(P, N, D, T, seed=seed)
        mask_o, ev_o = oracle_mod.feasible_mask(c["req"], c["unit_all"])
        mask, ev = engine.feasible_mask(engine.dev(c["req"], torch.float64), engine.dev(c["unit_all"], torch.float64))
        np.testing.assert_array_equal(to_np(mask), mask_o)
        assert int(to_np(ev)[0]) == ev_o
        used_o = np.zeros((N, D))
        oracle_mod.occupancy(c["row_ptr"], c["run_idx"], c["req_run"], used_o)
        d_used = engine.dev(np.zeros((N, D)), torch.float64)
        engine.occupancy(engine.dev(c["row_ptr"], torch.int64), engine.dev(c["run_idx"], torch.int32),
                         engine.dev(c["req_run"], torch.float64), d_used)
        np.testing.assert_array_equal(bits(to_np(d_used)), bits(used_o))


@pytest.mark.parametrize("N,D,T,seed,any_pending", [(100, 4, 1, 0, False), (5000, 4, 3, 1, True),
                                                     (20000, 8, 8, 2, False), (4097, 8, 5, 3, False)])
def test_node_states_and_maintain(engine, oracle_mod, N, D, T, seed, any_pending):
    c = syn.make_cluster(10, N, D, T, seed=seed, free_frac=0.6)
    thr = np.array([60, 300, 900, 1800, 3600, 7200, 21600, 86400], dtype=np.int64)
    st_o = oracle_mod.node_states(c["row_ptr"], c["run_idx"], c["req_run"], c["flags_run"], c["cap_type"],
                                  c["node_type"], c["node_flags"], c["node_age"], any_pending, thr)
    st = engine.node_states(engine.dev(c["row_ptr"], torch.int64), engine.dev(c["run_idx"], torch.int32),
                            engine.dev(c["req_run"], torch.float64), engine.dev(c["flags_run"], torch.uint8),
                            engine.dev(c["cap_type"], torch.float64), engine.dev(c["node_type"], torch.int32),
                            engine.dev(c["node_flags"], torch.uint8), engine.dev(c["node_age"], torch.int64),
                            any_pending, thr)
    np.testing.assert_array_equal(to_np(st), st_o)
    assert len(np.unique(st_o)) >= 3  # the generator exercises several states
    rng = np.random.default_rng(seed)
    for dry_run in (True, False):
        for budget in (np.zeros(T, np.int64), rng.integers(-2, 6, size=T).astype(np.int64),
                       np.full(T, 10 ** 6, np.int64)):
            scal = (rng.random(T) < 0.8).astype(np.uint8)
            s_o, a_o = oracle_mod.maintain_actions(st_o[3], c["node_pool"], budget, scal, dry_run)
            s_g, a_g = engine.maintain_actions(engine.dev(st_o[3].copy(), torch.uint8),
                                               engine.dev(c["node_pool"], torch.int32), budget, scal, dry_run)
            np.testing.assert_array_equal(to_np(s_g), s_o)
            np.testing.assert_array_equal(to_np(a_g), a_o)


def oracle_scale_up(oracle_mod, c, used):
    mask, ev0 = oracle_mod.feasible_mask(c["req"], c["unit_all"])
    feas_idx = np.nonzero(mask)[0]
    placed_f, ev1 = oracle_mod.first_fit_nodes(c["req"][feas_idx], c["cap_type"], c["node_type"], used)
    placed = np.where(mask.astype(bool), -1, -2).astype(np.int32)
    placed[feas_idx] = placed_f
    pend_idx = feas_idx[placed_f < 0]
    T = c["unit_ordered"].shape[0]
    res = {"feasible": mask, "placed": placed, "acc_pool": np.full(c["P"], -1, np.int32),
           "new_size": c["pool_actual"].astype(np.int64), "units_needed": np.full(T, -1, np.int64),
           "bins_opened": np.zeros(T, np.int64), "n_to_schedule": len(feas_idx), "n_pending": len(pend_idx),
           "num_unaccounted": 0, "decisions": ev0 + ev1}
    if len(pend_idx):
        f = oracle_mod.fulfill_pending(c["req"][pend_idx], len(pend_idx), c["unit_ordered"], c["pool_actual"],
                                       c["pool_max"], c["pool_ignored"], c["over_provision"])
        res["acc_pool"][pend_idx] = f["acc_pool"]
        for k in ("new_size", "units_needed", "bins_opened", "num_unaccounted"):
            res[k] = f[k]
        res["decisions"] += f["evals"]
    return res


@pytest.mark.parametrize("P,N,D,T,seed,max_size", [(16, 4, 4, 1, 0, None), (3000, 300, 4, 1, 1, None),
                                                    (20000, 2000, 8, 8, 2, None), (5000, 400, 4, 2, 3, 120)])
def test_scale_up_fused_and_host_entry(engine, oracle_mod, P, N, D, T, seed, max_size):
    c = syn.make_cluster(P, N, D, T, seed=seed, max_size=max_size, over_provision=seed % 2)
    used0 = syn.initial_used(c)
    used_o = used0.copy()
    o = oracle_scale_up(oracle_mod, c, used_o)
    # device-buffer entry
    d_used = engine.dev(used0, torch.float64)
    g = engine.scale_up(engine.dev(c["req"], torch.float64), c["unit_all"], c["unit_ordered"], c["pool_actual"],
                        c["pool_max"], c["pool_ignored"], c["over_provision"],
                        engine.dev(c["cap_type"], torch.float64), engine.dev(c["node_type"], torch.int32), d_used)
    torch.cuda.synchronize()
    # host-buffer entry (the plugin call)
    used_h = used0.copy()
    h = engine.scale_up_host(c["req"], c["unit_all"], c["unit_ordered"], c["pool_actual"], c["pool_max"],
                             c["pool_ignored"], c["over_provision"], c["cap_type"], c["node_type"], used_h)
    for res, used in ((g, to_np(d_used)), (h, used_h)):
        for k in ("feasible", "placed", "acc_pool"):
            v = res[k] if isinstance(res[k], np.ndarray) else to_np(res[k])
            np.testing.assert_array_equal(v, o[k], err_msg=k)
        for k in ("new_size", "units_needed", "bins_opened"):
            np.testing.assert_array_equal(res[k], o[k], err_msg=k)
        for k in ("n_to_schedule", "n_pending", "num_unaccounted", "decisions"):
            assert res[k] == o[k], k
        np.testing.assert_array_equal(bits(used), bits(used_o))


def test_maintain_host_entry(engine, oracle_mod):
    c = syn.make_cluster(10, 6000, 4, 3, seed=4, free_frac=0.5)
    budget = np.array([3, 0, 50], np.int64)
    scal = np.array([1, 1, 0], np.uint8)
    for dry_run in (True, False):
        st_o = oracle_mod.node_states(c["row_ptr"], c["run_idx"], c["req_run"], c["flags_run"], c["cap_type"],
                                      c["node_type"], c["node_flags"], c["node_age"], False, [1800])[0]
        s_o, a_o = oracle_mod.maintain_actions(st_o, c["node_pool"], budget, scal, dry_run)
        s, a = engine.maintain_host(c["row_ptr"], c["run_idx"], c["req_run"], c["flags_run"], c["cap_type"],
                                    c["node_type"], c["node_flags"], c["node_age"], c["node_pool"], False, 1800,
                                    budget, scal, dry_run)
        np.testing.assert_array_equal(s, s_o)
        np.testing.assert_array_equal(a, a_o)


def test_domain_errors(engine):
    from kubernetes_acs_engine_autoscaler_b200._native import AcsfitError
    req = np.ones((10, 4))
    req[3, 1] = -1.0
    cap = np.ones((1, 4)) * 4
    with pytest.raises(AcsfitError):
        engine.first_fit_nodes(engine.dev(req, torch.float64), None, engine.dev(cap, torch.float64),
                               engine.dev(np.zeros(2, np.int32), torch.int32), engine.dev(np.zeros((2, 4)), torch.float64))
    with pytest.raises(AcsfitError):  # D not in {2,4,8,16}
        engine.first_fit_nodes(engine.dev(np.ones((10, 3)), torch.float64), None, engine.dev(np.ones((1, 3)), torch.float64),
                               engine.dev(np.zeros(2, np.int32), torch.int32), engine.dev(np.zeros((2, 3)), torch.float64))


@pytest.mark.parametrize("N,D,seed", [(1, 2, 0), (31, 4, 1), (33, 4, 2), (1000, 8, 3), (2049, 16, 4), (70000, 4, 5)])
def test_streaming_kernels_ragged_and_permuted(engine, oracle_mod, N, D, seed):
    """K1 / K6 with what the warp-streaming form has to get right: a CSR whose index list is a random
    permutation of the rows (scattered gathers), empty nodes, nodes whose slice spans many staging chunks,
    node counts that are not a multiple of 32 / of the per-warp range, and every padded dimension count."""
    rng = np.random.default_rng(100 + seed)
    counts = rng.poisson(6.0, size=N).astype(np.int64)
    counts[rng.random(N) < 0.3] = 0                      # runs of empty nodes
    heavy = rng.integers(0, N, size=max(1, N // 200))
    counts[heavy] = rng.integers(300, 1500, size=len(heavy))  # slices much longer than a staging chunk
    row_ptr = np.zeros(N + 1, dtype=np.int64)
    np.cumsum(counts, out=row_ptr[1:])
    R = int(row_ptr[-1])
    Rtab = R + 17                                        # the row table is larger than the CSR and permuted
    run_idx = rng.permutation(Rtab)[:R].astype(np.int32)
    req_run = np.zeros((Rtab, D))
    req_run[:, 0] = rng.integers(1, 400, size=Rtab).astype(np.float64) * 1e-3
    req_run[:, 1] = rng.integers(1, 64, size=Rtab).astype(np.float64) * float(2 ** 20)
    if D > 2:
        req_run[:, 2:] = rng.integers(0, 3, size=(Rtab, D - 2)).astype(np.float64) * (rng.random((Rtab, D - 2)) < 0.3)
    flags_run = rng.integers(0, 4, size=Rtab).astype(np.uint8)
    T = 3
    cap_type = np.zeros((T, D))
    cap_type[:, 0] = [2.0, 4.0, 8.0]
    cap_type[:, 1] = [7e9, 14e9, 28e9]
    cap_type[:, 2:] = 110.0
    node_type = rng.integers(0, T, size=N).astype(np.int32)
    node_flags = (rng.random(N) < 0.1).astype(np.uint8)
    node_age = rng.integers(0, 86400, size=N).astype(np.int64)
    thr = np.array([60, 900, 3600, 86400], dtype=np.int64)

    used0 = rng.integers(0, 5, size=(N, D)).astype(np.float64) * 0.125
    used_o = used0.copy()
    oracle_mod.occupancy(row_ptr, run_idx, req_run, used_o)
    d_used = engine.dev(used0, torch.float64)
    engine.occupancy(engine.dev(row_ptr, torch.int64), engine.dev(run_idx, torch.int32), engine.dev(req_run, torch.float64), d_used)
    np.testing.assert_array_equal(bits(to_np(d_used)), bits(used_o))

    for any_pending in (False, True):
        st_o = oracle_mod.node_states(row_ptr, run_idx, req_run, flags_run, cap_type, node_type, node_flags, node_age,
                                      any_pending, thr)
        st = engine.node_states(engine.dev(row_ptr, torch.int64), engine.dev(run_idx, torch.int32),
                                engine.dev(req_run, torch.float64), engine.dev(flags_run, torch.uint8),
                                engine.dev(cap_type, torch.float64), engine.dev(node_type, torch.int32),
                                engine.dev(node_flags, torch.uint8), engine.dev(node_age, torch.int64), any_pending, thr)
        np.testing.assert_array_equal(to_np(st), st_o)


def test_chained_launch_timeout_falls_back_unchained(engine, oracle_mod):
    """a watchdog abort of the chained nodes+bins launch (possible when another context holds SMs) must not fail
    the tick: the node state is restored and the tick redone unchained, with the same result"""
    c = syn.make_cluster(20000, 2000, 4, 1, seed=31)
    used0 = syn.initial_used(c)
    args = (c["unit_all"], c["unit_ordered"], c["pool_actual"], c["pool_max"], c["pool_ignored"], c["over_provision"])
    f64, i32 = torch.float64, torch.int32

    def tick():
        used = engine.dev(used0, f64)
        r = engine.scale_up(engine.dev(c["req"], f64), *args, engine.dev(c["cap_type"], f64), engine.dev(c["node_type"], i32), used)
        return r, to_np(used)
    try:
        r_ok, used_ok = tick()
        engine.set_knob("inject_chain_timeout", 1)
        r_fb, used_fb = tick()          # first attempt "times out", second runs unchained
    finally:
        engine.set_knob("inject_chain_timeout", 0)
        engine.set_knob("overlap", 1)   # the fallback switches chaining off for the ctx: restore for later tests
    np.testing.assert_array_equal(to_np(r_ok["placed"]), to_np(r_fb["placed"]))
    np.testing.assert_array_equal(bits(used_ok), bits(used_fb))
    np.testing.assert_array_equal(r_ok["new_size"], r_fb["new_size"])
    assert r_ok["decisions"] == r_fb["decisions"] and r_ok["n_pending"] == r_fb["n_pending"] > 0


@pytest.mark.parametrize("N,D,seed", [(1, 2, 0), (31, 4, 1), (33, 4, 2), (1000, 8, 3), (2049, 16, 4), (70000, 4, 5), (4097, 2, 6),
                                      (300, 8, 7)])
def test_streaming_kernels_contiguous_table(engine, oracle_mod, N, D, seed):
    """K1 / K6 on a CONTIGUOUS running-pod table (run_idx = NULL -> bulk copies + mbarrier, acsfit_stream.cuh):
    warp ranges that start at any entry offset (chunks are re-aligned to 16 entries), tables whose length is not a
    multiple of 16 (the last flag bytes take the plain-load path), empty nodes, slices far longer than a chunk,
    every padded dimension count; the host-buffer calls take the same path."""
    rng = np.random.default_rng(300 + seed)
    counts = rng.poisson(7.0, size=N).astype(np.int64)
    counts[rng.random(N) < 0.25] = 0
    heavy = rng.integers(0, N, size=max(1, N // 150))
    counts[heavy] = rng.integers(200, 1200, size=len(heavy))
    if counts.sum() % 16 == 0:
        counts[0] += 3                                   # the table length is deliberately not a multiple of 16
    row_ptr = np.zeros(N + 1, dtype=np.int64)
    np.cumsum(counts, out=row_ptr[1:])
    R = int(row_ptr[-1])
    req_run = np.zeros((R, D))
    req_run[:, 0] = rng.integers(1, 400, size=R).astype(np.float64) * 1e-3
    req_run[:, 1] = rng.integers(1, 64, size=R).astype(np.float64) * float(2 ** 20)
    if D > 2:
        req_run[:, 2:] = rng.integers(0, 3, size=(R, D - 2)).astype(np.float64) * (rng.random((R, D - 2)) < 0.3)
    flags_run = rng.integers(0, 4, size=R).astype(np.uint8)
    T = 3
    cap_type = np.zeros((T, D))
    cap_type[:, 0] = [2.0, 4.0, 8.0]
    cap_type[:, 1] = [7e9, 14e9, 28e9]
    cap_type[:, 2:] = 110.0
    node_type = rng.integers(0, T, size=N).astype(np.int32)
    node_flags = (rng.random(N) < 0.1).astype(np.uint8)
    node_age = rng.integers(0, 86400, size=N).astype(np.int64)
    thr = np.array([60, 900, 3600, 86400], dtype=np.int64)
    ident = np.arange(R, dtype=np.int32)

    used0 = rng.integers(0, 5, size=(N, D)).astype(np.float64) * 0.125
    used_o = used0.copy()
    oracle_mod.occupancy(row_ptr, ident, req_run, used_o)
    d_used = engine.dev(used0, torch.float64)
    engine.occupancy(engine.dev(row_ptr, torch.int64), None, engine.dev(req_run, torch.float64), d_used)
    np.testing.assert_array_equal(bits(to_np(d_used)), bits(used_o))
    used_h = used0.copy()
    engine.occupancy_host(row_ptr, None, req_run, used_h)
    np.testing.assert_array_equal(bits(used_h), bits(used_o))

    for any_pending in (False, True):
        st_o = oracle_mod.node_states(row_ptr, ident, req_run, flags_run, cap_type, node_type, node_flags, node_age,
                                      any_pending, thr)
        st = engine.node_states(engine.dev(row_ptr, torch.int64), None, engine.dev(req_run, torch.float64),
                                engine.dev(flags_run, torch.uint8), engine.dev(cap_type, torch.float64),
                                engine.dev(node_type, torch.int32), engine.dev(node_flags, torch.uint8),
                                engine.dev(node_age, torch.int64), any_pending, thr)
        np.testing.assert_array_equal(to_np(st), st_o)
    node_pool = node_type.copy()
    budget = np.array([2, 0, 7], dtype=np.int64)
    s_o, a_o = oracle_mod.maintain_actions(st_o[1].copy(), node_pool, budget, np.ones(T, np.uint8), True)
    s_h, a_h = engine.maintain_host(row_ptr, None, req_run, flags_run, cap_type, node_type, node_flags, node_age,
                                    node_pool, True, 900, budget, np.ones(T, np.uint8), True)
    np.testing.assert_array_equal(s_h, s_o)
    np.testing.assert_array_equal(a_h, a_o)
