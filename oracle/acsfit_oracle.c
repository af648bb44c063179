/*
 * acsfit_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, single-threaded restatement of the reference autoscaler's per-tick
 * decision path on a dense float64 layout ("absent resource key == 0.0").
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product (libacsfit.so)
 * never links or calls it.
 *
 * Parity pin: oracle/make_golden.py runs the UNMODIFIED reference
 * (/root/reference/autoscaler, imported under oracle/ref_shim.py) on seeded
 * cluster states and on the reference's own known-answer tests
 * (test/test_cluster.py:56-73, test/test_scaler.py:53-77) and commits the
 * results under tests/golden/; tests/test_host_golden_cpu.py (whole ticks replayed with this
 * library as the engine) and tests/test_gpu_parity.py (CUDA vs this library) check every
 * function below against those vectors bit for bit.
 *
 * Every arithmetic expression keeps the reference's operation ORDER:
 *   node fit   : cap - (used + req) >= 0      autoscaler/kube.py:173-176, :203-213, :247-249
 *   bin fit    : remaining - req   >= 0       autoscaler/scaler.py:134,139-140
 *   under-util : 0.3*cap - util    >= 0       autoscaler/scaler.py:85-87   (mul THEN sub: no FMA)
 * Build with -ffp-contract=off (oracle/Makefile) so gcc cannot fuse the last one.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

/* ClusterNodeState codes, in the order autoscaler/scaler.py:19-29 lists them */
enum {
    ST_INSTANCE_TERMINATED = 0, ST_POD_PENDING = 1, ST_GRACE_PERIOD = 2, ST_SPARE_AGENT = 3,
    ST_IDLE_SCHEDULABLE = 4, ST_IDLE_UNSCHEDULABLE = 5, ST_BUSY_UNSCHEDULABLE = 6, ST_BUSY = 7,
    ST_UNDER_UTILIZED_DRAINABLE = 8, ST_UNDER_UTILIZED_UNDRAINABLE = 9
};
/* maintain() action codes (autoscaler/engine_scaler.py:149-182) */
enum { ACT_NONE = 0, ACT_CORDON_DRAIN = 1, ACT_CORDON = 2, ACT_UNCORDON = 3, ACT_SCALE_IN = 4 };

/* per running-pod flag bits (host-computed from annotations/labels/name, kube.py:51-71) */
#define PODF_BUSY        1u  /* not is_mirrored() and 'kube-proxy' not in name   scaler.py:76 */
#define PODF_UNDRAINABLE 2u  /* not (is_drainable() or 'kube-proxy' in name)     scaler.py:82-83 */
/* per node flag bits */
#define NODEF_UNSCHEDULABLE 1u

/* (a - b).possible for one pod row against one capacity-like row: kube.py:209-213,:247-249 */
static int sub_possible(const double *a, const double *b, int D)
{
    for (int d = 0; d < D; ++d) {
        double diff = a[d] - b[d];
        if (!(diff >= 0)) return 0;
    }
    return 1;
}

/* KubeNode.can_fit: capacity - (used_capacity + resources) possible   kube.py:173-176 */
static int node_can_fit(const double *cap, const double *used, const double *req, int D)
{
    for (int d = 0; d < D; ++d) {
        double s = used[d] + req[d];
        double left = cap[d] - s;
        if (!(left >= 0)) return 0;
    }
    return 1;
}

/*
 * capacity.is_possible over every agent pool (capacity.py:24-32) as used by
 * Cluster.get_pods_to_schedule (cluster.py:217-240).  unit[T][D] holds
 * RESOURCE_SPEC[pool.instance_type] for every pool (ignored pools included).
 * out_mask[p] = 1 when the pod fits on at least one pool's unit.
 * Returns the number of (unit - req).possible evaluations the reference performs
 * (it stops at the first pool that fits).
 */
ORACLE_API uint64_t oracle_feasible_mask(const double *req, int64_t P, int D,
                                         const double *unit, int T, uint8_t *out_mask)
{
    uint64_t evals = 0;
    for (int64_t p = 0; p < P; ++p) {
        uint8_t ok = 0;
        for (int t = 0; t < T; ++t) {
            ++evals;
            if (sub_possible(unit + (size_t)t * D, req + (size_t)p * D, D)) { ok = 1; break; }
        }
        out_mask[p] = ok;
    }
    return evals;
}

/*
 * Occupancy accumulation, cluster.py:165-168 + KubeNode.count_pod kube.py:169-171:
 * for every node, used = used + req for each running/assigned pod whose
 * node_name equals the node's name, in POD-LIST order.  The pods of node n are
 * run_idx[row_ptr[n] .. row_ptr[n+1]) (indices into req_run, ascending).
 */
ORACLE_API void oracle_occupancy(const int64_t *row_ptr, const int32_t *run_idx,
                                 const double *req_run, int64_t N, int D, double *used)
{
    for (int64_t n = 0; n < N; ++n)
        for (int64_t k = row_ptr[n]; k < row_ptr[n + 1]; ++k) {
            const double *r = req_run + (size_t)run_idx[k] * D;
            for (int d = 0; d < D; ++d) {
                double s = used[(size_t)n * D + d] + r[d];
                used[(size_t)n * D + d] = s;
            }
        }
}

/*
 * Cluster.get_pending_pods, cluster.py:184-204: sequential first-fit of the
 * pods (in list order) over the nodes (in list order); a hit mutates
 * used[node] += req (count_pod).  placed[p] = node index or -1 (pending).
 * Returns the number of KubeNode.can_fit calls (the credited "decisions").
 */
ORACLE_API uint64_t oracle_first_fit_nodes(const double *req, int64_t P, int D,
                                           const double *cap_type, const int32_t *node_type,
                                           double *used, int64_t N, int32_t *placed)
{
    uint64_t calls = 0;
    for (int64_t p = 0; p < P; ++p) {
        const double *r = req + (size_t)p * D;
        int64_t hit = -1;
        for (int64_t n = 0; n < N; ++n) {
            ++calls;
            if (node_can_fit(cap_type + (size_t)node_type[n] * D, used + (size_t)n * D, r, D)) {
                hit = n;
                break;
            }
        }
        placed[p] = (int32_t)hit;
        if (hit >= 0)
            for (int d = 0; d < D; ++d) {
                double s = used[(size_t)hit * D + d] + r[d];
                used[(size_t)hit * D + d] = s;
            }
    }
    return calls;
}

/*
 * Scaler.fulfill_pending, scaler.py:117-177 (up to, not including, the
 * raise / scale_pools hand-off at :179-184, which is host logic).
 *
 *  req[Pp][D]     the UNIQUE pending pods in dict order (scaler.py:119: duplicate
 *                 uids collapse to their first occurrence)
 *  num_listed     len(pods) including duplicates (scaler.py:120)
 *  unit[T][D]     pool unit capacities, pools ALREADY in visiting order
 *                 (capacity.order_by_cost_asc, capacity.py:34-36)
 *  pool_actual / pool_max / pool_ignored   agent_pool.py:18-23, scaler.py:128
 * out:
 *  new_size[T]        new_pool_sizes values (scaler.py:125,167)
 *  units_needed[T]    len(bins)+over_provision, or -1 when the pool was skipped (:128-129)
 *  bins_opened[T]     len(new_instance_resources), 0 when skipped
 *  acc_pool[Pp]       visiting index of the pool that accounted the pod, else -1
 *  bin_of[Pp]         bin index inside the LAST pool that packed the pod, else -1
 *  out_unaccounted    num_unaccounted after the loop (scaler.py:179)
 * Returns the number of `.possible` evaluations (pool gate :134 + bin tests :139).
 */
ORACLE_API uint64_t oracle_fulfill_pending(const double *req, int64_t Pp, int64_t num_listed, int D,
                                           const double *unit, const int32_t *pool_actual,
                                           const int32_t *pool_max, const uint8_t *pool_ignored,
                                           int T, int64_t over_provision,
                                           int64_t *new_size, int64_t *units_needed,
                                           int64_t *bins_opened, int32_t *acc_pool,
                                           int32_t *bin_of, int64_t *out_unaccounted)
{
    uint64_t evals = 0;
    int64_t num_unaccounted = num_listed;
    double *bins = (double *)malloc(sizeof(double) * (size_t)(Pp > 0 ? Pp : 1) * D);
    int32_t *cur_bin = (int32_t *)malloc(sizeof(int32_t) * (size_t)(Pp > 0 ? Pp : 1));
    for (int64_t p = 0; p < Pp; ++p) { acc_pool[p] = -1; bin_of[p] = -1; }

    for (int t = 0; t < T; ++t) {
        const double *u = unit + (size_t)t * D;
        new_size[t] = pool_actual[t];
        units_needed[t] = -1;
        bins_opened[t] = 0;
        if (pool_ignored[t] || !num_unaccounted) continue;

        int64_t nb = 0;
        for (int64_t p = 0; p < Pp; ++p) {
            cur_bin[p] = -1;
            if (acc_pool[p] >= 0) continue;
            const double *r = req + (size_t)p * D;
            ++evals;
            if (!sub_possible(u, r, D)) continue;
            int64_t hit = -1;
            for (int64_t i = 0; i < nb; ++i) {
                ++evals;
                if (sub_possible(bins + (size_t)i * D, r, D)) { hit = i; break; }
            }
            if (hit < 0) {
                hit = nb++;
                for (int d = 0; d < D; ++d) bins[(size_t)hit * D + d] = u[d];
            }
            for (int d = 0; d < D; ++d) {
                double left = bins[(size_t)hit * D + d] - r[d];
                bins[(size_t)hit * D + d] = left;
            }
            cur_bin[p] = (int32_t)hit;
            bin_of[p] = (int32_t)hit;
        }
        int64_t needed = nb + over_provision;
        int64_t room = (int64_t)pool_max[t] - (int64_t)pool_actual[t];
        int64_t unavailable = needed - room > 0 ? needed - room : 0;
        int64_t requested = needed - unavailable;
        units_needed[t] = needed;
        bins_opened[t] = nb;
        new_size[t] = (int64_t)pool_actual[t] + requested;
        int64_t take = nb < requested ? nb : requested; /* range(min(len(bins), requested)) */
        for (int64_t p = 0; p < Pp; ++p)
            if (cur_bin[p] >= 0 && cur_bin[p] < take) {
                acc_pool[p] = t;
                --num_unaccounted;
            }
    }
    free(bins);
    free(cur_bin);
    *out_unaccounted = num_unaccounted;
    return evals;
}

/*
 * Scaler.get_node_state, scaler.py:61-114, for every node, for S idle
 * thresholds at once (out_state[s][n]).  The node's pods are
 * run_idx[row_ptr[n]..row_ptr[n+1]) in pods_by_node order (engine_scaler.py:129-131).
 */
ORACLE_API void oracle_node_states(const int64_t *row_ptr, const int32_t *run_idx,
                                   const double *req_run, const uint8_t *flags_run,
                                   const double *cap_type, const int32_t *node_type,
                                   const uint8_t *node_flags, const int64_t *node_age,
                                   int64_t N, int D, int any_pending,
                                   const int64_t *idle_threshold, int S, uint8_t *out_state)
{
    double *util = (double *)malloc(sizeof(double) * (size_t)D);
    for (int64_t n = 0; n < N; ++n) {
        int busy = 0, undrainable = 0;
        for (int d = 0; d < D; ++d) util[d] = 0.0;
        for (int64_t k = row_ptr[n]; k < row_ptr[n + 1]; ++k) {
            int32_t j = run_idx[k];
            if (flags_run[j] & PODF_UNDRAINABLE) undrainable = 1;
            if (flags_run[j] & PODF_BUSY) {
                busy = 1;
                for (int d = 0; d < D; ++d) {
                    double s = util[d] + req_run[(size_t)j * D + d];
                    util[d] = s;
                }
            }
        }
        const double *cap = cap_type + (size_t)node_type[n] * D;
        int under = 1;
        for (int d = 0; d < D; ++d) {
            double thr = cap[d] * 0.3;       /* UTIL_THRESHOLD * capacity  (KubeResource.__rmul__) */
            double left = thr - util[d];
            if (!(left >= 0)) { under = 0; break; }
        }
        int unsched = (node_flags[n] & NODEF_UNSCHEDULABLE) != 0;
        for (int s = 0; s < S; ++s) {
            uint8_t st;
            if (busy && !under) st = unsched ? ST_BUSY_UNSCHEDULABLE : ST_BUSY;
            else if (any_pending && !unsched) st = ST_POD_PENDING;
            else if (node_age[n] <= idle_threshold[s] && !unsched) st = ST_GRACE_PERIOD;
            else if (under && (busy || !unsched))
                st = undrainable ? ST_UNDER_UTILIZED_UNDRAINABLE : ST_UNDER_UTILIZED_DRAINABLE;
            else st = unsched ? ST_IDLE_UNSCHEDULABLE : ST_IDLE_SCHEDULABLE;
            out_state[(size_t)s * N + n] = st;
        }
    }
    free(util);
}

/*
 * EngineScaler.maintain's decision part, engine_scaler.py:133-182, for one
 * state vector.  node_pool[n] = index of the node's pool; nodes are visited pool
 * by pool in node-list order (pool.nodes, engine_scaler.py:43-45), only pools
 * with pool_scalable[t] != 0.  budget0[t] = actual - len(unschedulable) - spare
 * (:136).  The budget decrements only when not dry_run (:154-159).
 * io_state is rewritten in place with SPARE_AGENT substitutions (:142-144);
 * out_action gets the action code; nodes of ignored pools keep action NONE and
 * their state is reported as 255 (never evaluated by the reference).
 */
ORACLE_API void oracle_maintain_actions(uint8_t *io_state, const int32_t *node_pool, int64_t N,
                                        const int64_t *budget0, const uint8_t *pool_scalable, int T,
                                        int dry_run, uint8_t *out_action)
{
    int64_t *budget = (int64_t *)malloc(sizeof(int64_t) * (size_t)(T > 0 ? T : 1));
    for (int t = 0; t < T; ++t) budget[t] = budget0[t];
    for (int64_t n = 0; n < N; ++n) {
        int t = node_pool[n];
        out_action[n] = ACT_NONE;
        if (t < 0 || t >= T || !pool_scalable[t]) { io_state[n] = 255; continue; }
        uint8_t st = io_state[n];
        if (st == ST_UNDER_UTILIZED_DRAINABLE && budget[t] == 0) st = ST_SPARE_AGENT;
        io_state[n] = st;
        switch (st) {
        case ST_UNDER_UTILIZED_DRAINABLE:
            out_action[n] = ACT_CORDON_DRAIN;
            if (!dry_run) budget[t] -= 1;
            break;
        case ST_IDLE_SCHEDULABLE: out_action[n] = ACT_CORDON; break;
        case ST_BUSY_UNSCHEDULABLE: out_action[n] = ACT_UNCORDON; break;
        case ST_IDLE_UNSCHEDULABLE: out_action[n] = ACT_SCALE_IN; break;
        default: break;
        }
    }
    free(budget);
}

/* number of `KubeNode.can_fit` calls implied by a placement vector (for crediting) */
ORACLE_API uint64_t oracle_count_decisions(const int32_t *placed, int64_t P, int64_t N)
{
    uint64_t c = 0;
    for (int64_t p = 0; p < P; ++p) c += placed[p] >= 0 ? (uint64_t)placed[p] + 1 : (uint64_t)N;
    return c;
}
