"""TEST INFRASTRUCTURE: a stand-in for kubernetes_acs_engine_autoscaler_b200.engine.Engine that
answers every call with the plain-C oracle on CPU tensors.  It exists so that the HOST logic
(flattening, key-set bookkeeping, log text, error behaviour, adapter calls) can be tested in
`-m "not gpu"` runs; the product never imports it and has no CPU path of its own."""
import numpy as np
import torch

import oracle


class OracleEngine(object):
    device = torch.device("cpu")

    def dev(self, a, dtype):
        if isinstance(a, torch.Tensor):
            return a.to(dtype=dtype).contiguous().clone()
        return torch.from_numpy(np.ascontiguousarray(a)).to(dtype=dtype).contiguous().clone()

    def empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype)

    def feasible_mask(self, req, unit):
        mask, evals = oracle.feasible_mask(req.numpy(), unit.numpy())
        return torch.from_numpy(mask), torch.tensor([evals], dtype=torch.int64)

    def occupancy(self, row_ptr, run_idx, req_run, used):
        u = used.numpy()
        idx = np.arange(req_run.shape[0], dtype=np.int32) if run_idx is None else run_idx.numpy()  # None: contiguous table
        oracle.occupancy(row_ptr.numpy(), idx, req_run.numpy(), u)
        return used

    def first_fit_nodes(self, req, pod_idx, cap_type, node_type, used):
        rows = req.numpy() if pod_idx is None else req.numpy()[pod_idx.numpy()]
        u = used.numpy()
        placed, calls = oracle.first_fit_nodes(rows, cap_type.numpy(), node_type.numpy(), u)
        return torch.from_numpy(placed), torch.tensor([calls], dtype=torch.int64)

    def fulfill_pending(self, req, num_listed, unit, pool_actual, pool_max, pool_ignored, over_provision):
        r = oracle.fulfill_pending(req.numpy(), num_listed, np.asarray(unit, dtype=np.float64).reshape(-1, req.shape[1]),
                                   pool_actual, pool_max, pool_ignored, over_provision)
        r["acc_pool"] = torch.from_numpy(r["acc_pool"])
        r["bin_of"] = torch.from_numpy(r["bin_of"])
        return r

    def node_states(self, row_ptr, run_idx, req_run, flags_run, cap_type, node_type, node_flags, node_age,
                    any_pending, idle_thresholds):
        idx = np.arange(req_run.shape[0], dtype=np.int32) if run_idx is None else run_idx.numpy()
        st = oracle.node_states(row_ptr.numpy(), idx, req_run.numpy(), flags_run.numpy(), cap_type.numpy(),
                                node_type.numpy(), node_flags.numpy(), node_age.numpy(), any_pending, idle_thresholds)
        return torch.from_numpy(st)

    def maintain_actions(self, state, node_pool, budget0, pool_scalable, dry_run):
        s, a = oracle.maintain_actions(state.numpy(), node_pool.numpy(), budget0, pool_scalable, dry_run)
        return torch.from_numpy(s), torch.from_numpy(a)
