"""`-m "not gpu"`: the plain-C oracle against the mid-size ticks recorded from the unmodified reference
(tests/golden/mid_ticks.json.gz: 256-node stage crossings, several pod tiles, D = 8 / T = 8, the raise path)."""
import numpy as np
import pytest

import mid_golden


@pytest.mark.parametrize("name", mid_golden.case_ids())
def test_oracle_reproduces_reference_tick(oracle_mod, name):
    case = [c for c in mid_golden.load_cases() if c["name"] == name][0]
    c = mid_golden.cluster_of(case)
    from kubernetes_acs_engine_autoscaler_b200 import synthetic as syn
    used = syn.initial_used(c)
    mask, _ = oracle_mod.feasible_mask(c["req"], c["unit_all"])
    idx = np.nonzero(mask)[0]
    placed_f, _ = oracle_mod.first_fit_nodes(c["req"][idx], c["cap_type"], c["node_type"], used)
    placed = np.where(mask.astype(bool), -1, -2).astype(np.int32)
    placed[idx] = placed_f
    pend = idx[placed_f < 0]
    new_size, unacc = c["pool_actual"].astype(np.int64), 0
    if len(pend):
        f = oracle_mod.fulfill_pending(c["req"][pend], len(pend), c["unit_ordered"], c["pool_actual"], c["pool_max"],
                                       c["pool_ignored"], c["over_provision"])
        new_size, unacc = f["new_size"], f["num_unaccounted"]
    st = oracle_mod.node_states(c["row_ptr"], c["run_idx"], c["req_run"], c["flags_run"], c["cap_type"], c["node_type"],
                                c["node_flags"], c["node_age"], len(idx) > 0, [1800])[0]
    mid_golden.check_tick(case, placed, used, new_size, unacc, st)
