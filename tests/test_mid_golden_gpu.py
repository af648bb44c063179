"""`-m gpu`: the CUDA path (through the C ABI) against the mid-size ticks recorded from the unmodified reference,
in every schedule of the pipeline: default, forced scan-list pruning, no nodes->bins chaining (and, through the
`engine` fixture, both the packed-rank and the float64 form of the candidate scan)."""
import numpy as np
import pytest
import torch

import mid_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("schedule", ["default", "prune", "no_overlap", "min_stages_3"])
@pytest.mark.parametrize("name", mid_golden.case_ids())
def test_cuda_reproduces_reference_tick(engine, name, schedule):
    case = [c for c in mid_golden.load_cases() if c["name"] == name][0]
    c = mid_golden.cluster_of(case)
    engine.set_knob("prune", 1 if schedule == "prune" else -1)
    engine.set_knob("overlap", 0 if schedule == "no_overlap" else 1)
    engine.set_knob("min_stages", 3 if schedule == "min_stages_3" else 0)
    try:
        f64, i32, i64, u8 = torch.float64, torch.int32, torch.int64, torch.uint8
        used = engine.dev(np.zeros((c["N"], c["D"])), f64)
        engine.occupancy(engine.dev(c["row_ptr"], i64), engine.dev(c["run_idx"], i32), engine.dev(c["req_run"], f64), used)
        r = engine.scale_up(engine.dev(c["req"], f64), c["unit_all"], c["unit_ordered"], c["pool_actual"], c["pool_max"],
                            c["pool_ignored"], c["over_provision"], engine.dev(c["cap_type"], f64),
                            engine.dev(c["node_type"], i32), used)
        st = engine.node_states(engine.dev(c["row_ptr"], i64), engine.dev(c["run_idx"], i32), engine.dev(c["req_run"], f64),
                                engine.dev(c["flags_run"], u8), engine.dev(c["cap_type"], f64),
                                engine.dev(c["node_type"], i32), engine.dev(c["node_flags"], u8),
                                engine.dev(c["node_age"], i64), r["n_to_schedule"] > 0, [1800])
        mid_golden.check_tick(case, r["placed"].cpu().numpy(), used.cpu().numpy(), r["new_size"], r["num_unaccounted"],
                              st[0].cpu().numpy())
    finally:
        engine.set_knob("prune", -1)
        engine.set_knob("overlap", 1)
        engine.set_knob("min_stages", 0)
