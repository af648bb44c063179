/*
 * acsfit.h -- C ABI of libacsfit.so, the B200 (sm_100a) decision engine behind the
 * acs-engine autoscaler's per-tick hot path.
 *
 * The reference (wbuchwalter/Kubernetes-acs-engine-autoscaler) is pure Python and has no
 * FFI: the boundary it offers is a set of Python methods.  Each entry point below replaces
 * the arithmetic of one of those methods on a dense float64 snapshot ("absent resource key
 * == 0.0"); the Python classes of the same names in
 * kubernetes_acs_engine_autoscaler_b200/ translate between kube objects and these buffers
 * (INTEGRATION.md shows the ctypes binding a reference maintainer would add).
 *
 * Conventions
 *  - plain C: pointers and sizes only, no torch / CUDA types in the signatures.
 *    `acsfit_stream_t` is a `cudaStream_t` passed as `void*` (NULL = default stream).
 *  - every function returns an `acsfit_status` (0 = ok, < 0 = error); nothing throws
 *    across the boundary.  `acsfit_last_error(ctx)` gives the message.
 *  - pointers marked [dev] are device pointers owned by the caller (e.g.
 *    torch.Tensor.data_ptr()); the library never frees or retains them past the call.
 *    Pointers marked [host] are host pointers.  Small per-pool results come back through
 *    [host] pointers, so those calls synchronise the stream before returning.
 *  - matrices are row-major float64 with exactly D columns; D <= ACSFIT_MAX_DIMS.
 *  - one ctx per host thread (thread-compatible, not thread-safe); the ctx owns a reusable
 *    scratch arena so a steady-state tick performs no cudaMalloc.
 *  - input domain: requests are >= 0 and not NaN (reference: utils.py:33 regex admits no
 *    sign); capacity / unit rows are finite.  Violations return ACSFIT_E_DOMAIN.
 */
#ifndef ACSFIT_H
#define ACSFIT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ACSFIT_ABI_VERSION 2
#define ACSFIT_MAX_DIMS 16
#define ACSFIT_MAX_POOLS 256

#if defined(__GNUC__)
#define ACSFIT_API __attribute__((visibility("default")))
#else
#define ACSFIT_API
#endif

typedef int acsfit_status;
enum {
    ACSFIT_OK = 0,
    ACSFIT_E_INVALID = -1,  /* bad argument (NULL pointer, negative size, D too large ...) */
    ACSFIT_E_CUDA = -2,     /* a CUDA runtime call or kernel failed */
    ACSFIT_E_DOMAIN = -3,   /* input outside the reference's domain (negative / NaN request) */
    ACSFIT_E_TIMEOUT = -4,  /* device-side watchdog fired (inter-stage pipeline stalled) */
    ACSFIT_E_NOMEM = -5,
    ACSFIT_E_PEER = -6      /* cluster mode: a peer rank failed / did not arrive (in-stream barrier timed out) */
};

/* ClusterNodeState codes, in the order reference autoscaler/scaler.py:19-29 lists them */
enum {
    ACSFIT_ST_INSTANCE_TERMINATED = 0,
    ACSFIT_ST_POD_PENDING = 1,
    ACSFIT_ST_GRACE_PERIOD = 2,
    ACSFIT_ST_SPARE_AGENT = 3,
    ACSFIT_ST_IDLE_SCHEDULABLE = 4,
    ACSFIT_ST_IDLE_UNSCHEDULABLE = 5,
    ACSFIT_ST_BUSY_UNSCHEDULABLE = 6,
    ACSFIT_ST_BUSY = 7,
    ACSFIT_ST_UNDER_UTILIZED_DRAINABLE = 8,
    ACSFIT_ST_UNDER_UTILIZED_UNDRAINABLE = 9,
    ACSFIT_ST_NOT_EVALUATED = 255 /* node of an ignored pool: the reference never asks */
};
/* maintain() actions (reference autoscaler/engine_scaler.py:149-182) */
enum {
    ACSFIT_ACT_NONE = 0,
    ACSFIT_ACT_CORDON_DRAIN = 1, /* UNDER_UTILIZED_DRAINABLE :153-161 */
    ACSFIT_ACT_CORDON = 2,       /* IDLE_SCHEDULABLE         :162-166 */
    ACSFIT_ACT_UNCORDON = 3,     /* BUSY_UNSCHEDULABLE       :167-172 */
    ACSFIT_ACT_SCALE_IN = 4      /* IDLE_UNSCHEDULABLE       :173-177 */
};
/* per running-pod flag bits, computed by the host from annotations / labels / name */
#define ACSFIT_PODF_BUSY 1u        /* not is_mirrored() and 'kube-proxy' not in name  scaler.py:76    */
#define ACSFIT_PODF_UNDRAINABLE 2u /* not (is_drainable() or 'kube-proxy' in name)    scaler.py:82-83 */
/* per node flag bits */
#define ACSFIT_NODEF_UNSCHEDULABLE 1u /* spec.unschedulable  kube.py:105 */

typedef struct acsfit_ctx acsfit_ctx;
typedef void *acsfit_stream_t;

ACSFIT_API int acsfit_abi_version(void);
ACSFIT_API const char *acsfit_last_error(const acsfit_ctx *ctx);

/* `device` is the CUDA device ordinal.  Fails (ACSFIT_E_CUDA) when no sm_100 device is there:
 * there is no CPU fallback. */
ACSFIT_API acsfit_status acsfit_ctx_create(int device, acsfit_ctx **out_ctx);
ACSFIT_API acsfit_status acsfit_ctx_destroy(acsfit_ctx *ctx);
/* pipeline tuning knobs (0 = library default): stages the node/bin axis is cut into at least,
 * and the watchdog in milliseconds.
 * Developer knobs read from the environment at acsfit_ctx_create (none changes a result, only the schedule):
 *   ACSFIT_OVERLAP=0        do not chain the first bin pass behind the node pass on a second stream
 *   ACSFIT_PRUNE=0|1        never / always use the scan-list-pruning instantiation of the node pass
 *                           (default: only when the pass has more stages than the GPU holds at once)
 *   ACSFIT_MIN_STAGES=n     same as acsfit_ctx_configure(min_stages) (default: one stage per SM; measured best
 *                           for the chained c2 tick as well: 148 -> 10.86 ms, 79 -> 11.13, 40 -> 11.4)
 *   ACSFIT_SMEM_FLOOR_KB=n  request at least n KB of shared memory per stage CTA (limits CTAs per SM)
 *   ACSFIT_STREAM_BYTES=n   staging bytes per warp of the K1/K6 streaming kernels (2048 / 4096 / 8192)
 *   ACSFIT_RANKS=0          never use the packed-rank scan (float64 compares instead; same results) */
ACSFIT_API acsfit_status acsfit_ctx_configure(acsfit_ctx *ctx, int min_stages, int watchdog_ms);
/* the same developer knobs by name, after ctx creation: "ranks" (0: float64 compare scan, 1: packed-rank scan when
 * the tick's request table allows it -- the default), "stream" (1: the barrier-free streaming form of the pipeline
 * kernel, csrc/acsfit_stream_ff.cuh, where it applies; env ACSFIT_STREAM), "prune", "overlap", "min_stages", "inject_chain_timeout" (test hook), "cluster_blocks" (cluster
 * mode: cut the pod list of the node pass into this many blocks; 0 = automatic).  None changes a result. */
ACSFIT_API acsfit_status acsfit_ctx_set_knob(acsfit_ctx *ctx, const char *name, int value);
/* when enabled, the first-fit / bin-pack pipeline launches are bracketed with CUDA events on the
 * launching stream (read back with acsfit_last_pipeline_stats) */
ACSFIT_API acsfit_status acsfit_ctx_set_timing(acsfit_ctx *ctx, int enabled);

/*
 * K0 feasible_mask -- capacity.is_possible over all agent pools, as used by
 * Cluster.get_pods_to_schedule (reference autoscaler/capacity.py:24-32, cluster.py:217-240).
 *   req   [dev] P x D   pending-unassigned pods, list order
 *   unit  [dev] T x D   RESOURCE_SPEC[pool.instance_type] per agent pool (ignored pools included)
 *   out_mask  [dev] P bytes: 1 when (unit[t] - req).possible for some t
 *   out_evals [dev] one uint64 (may be NULL): number of `.possible` evaluations the reference
 *             performs (it stops at the first pool that fits)
 */
ACSFIT_API acsfit_status acsfit_feasible_mask(acsfit_ctx *ctx, const double *req, int64_t P, int D,
                                   const double *unit, int T, uint8_t *out_mask,
                                   uint64_t *out_evals, acsfit_stream_t stream);

/*
 * K1 occupancy -- the node.count_pod loop of Cluster.loop_logic (reference cluster.py:165-168,
 * kube.py:169-171): used[n] = used[n] + req_run[j] for the node's running/assigned pods in
 * POD-LIST order (left-to-right float64 sum, no tree reduction).
 *   row_ptr [dev] N+1 int64, run_idx [dev] row_ptr[N] int32 indices into req_run (ascending
 *   inside each node), req_run [dev] R x D, used_inout [dev] N x D (zero = KubeResource()).
 *   run_idx may be NULL (here, in acsfit_node_states and in the *_host forms): the table is then CONTIGUOUS, the
 *   pods of node n are rows row_ptr[n] .. row_ptr[n+1] of req_run / flags_run.  That is the layout the host
 *   layer produces, and it lets K1 / K6 stream the table with bulk copies (csrc/acsfit_stream.cuh).
 */
ACSFIT_API acsfit_status acsfit_occupancy(acsfit_ctx *ctx, const int64_t *row_ptr, const int32_t *run_idx,
                               const double *req_run, int64_t N, int D, double *used_inout,
                               acsfit_stream_t stream);

/*
 * K2+K3 first_fit_nodes -- Cluster.get_pending_pods (reference cluster.py:184-204) with
 * KubeNode.can_fit (kube.py:173-176): sequential first-fit of pods (list order) over nodes
 * (list order); a hit mutates used[node] += req.  Exact and order-preserving.
 *   req        [dev] req_rows x D     (D must be 2, 4, 8 or 16: zero columns are neutral, because
 *                                      an absent resource key is 0.0 on both sides of the test)
 *   pod_idx    [dev] P int32 row numbers into req, the pods to schedule in order; NULL = rows 0..P-1
 *   cap_type   [dev] K x D, node_type [dev] N int32 (row of cap_type per node)
 *   used_inout [dev] N x D
 *   out_placed [dev] P int32: node index, or -1 = pending
 *   out_decisions [dev] one uint64 (may be NULL): number of KubeNode.can_fit calls the
 *                 reference makes = sum(placed >= 0 ? placed + 1 : N)
 */
ACSFIT_API acsfit_status acsfit_first_fit_nodes(acsfit_ctx *ctx, const double *req, int64_t req_rows,
                                     const int32_t *pod_idx, int64_t P, int D, const double *cap_type,
                                     const int32_t *node_type, double *used_inout, int64_t N,
                                     int32_t *out_placed, uint64_t *out_decisions,
                                     acsfit_stream_t stream);

/*
 * K4+K5 fulfill_pending -- Scaler.fulfill_pending (reference scaler.py:117-177): per pool in
 * cost order, first-fit of the still-unaccounted eligible pods into new-instance bins, the
 * pool-size clamp arithmetic and the accounting of the first min(len(bins), units_requested)
 * bins.  The raise / scale_pools hand-off (:179-184) stays in the host layer.
 *   req [dev] Pp x D      UNIQUE pending pods in dict order (duplicate uids collapsed, :119);
 *                         D must be 2, 4, 8 or 16 (zero columns are neutral)
 *   num_listed            len(pods) including duplicates (:120)
 *   unit [host] T x D     pool unit capacities, pools ALREADY in visiting order (capacity.py:34-36)
 *   pool_actual/pool_max [host] T int32, pool_ignored [host] T bytes
 *   out_new_size/out_units_needed/out_bins_opened [host] T int64 (units_needed = -1: pool skipped)
 *   out_acc_pool [dev] Pp int32 visiting index of the pool that accounted the pod, -1 = none
 *   out_bin_of   [dev] Pp int32 bin index inside the last pool that packed the pod, -1 = none
 *   out_unaccounted [host] one int64 (num_unaccounted at :179)
 *   out_evals [host] one uint64 (may be NULL): `.possible` evaluations (pool gate :134 + bins :139)
 */
ACSFIT_API acsfit_status acsfit_fulfill_pending(acsfit_ctx *ctx, const double *req, int64_t Pp,
                                     int64_t num_listed, int D, const double *unit,
                                     const int32_t *pool_actual, const int32_t *pool_max,
                                     const uint8_t *pool_ignored, int T, int64_t over_provision,
                                     int64_t *out_new_size, int64_t *out_units_needed,
                                     int64_t *out_bins_opened, int32_t *out_acc_pool,
                                     int32_t *out_bin_of, int64_t *out_unaccounted,
                                     uint64_t *out_evals, acsfit_stream_t stream);

/*
 * K6 node_states -- Scaler.get_node_state (reference scaler.py:61-114) for every node and S
 * idle thresholds at once.  The node's pods are run_idx[row_ptr[n]..row_ptr[n+1]) in
 * pods_by_node order (engine_scaler.py:129-131).
 *   flags_run [dev] R bytes (ACSFIT_PODF_*), node_flags [dev] N bytes (ACSFIT_NODEF_*),
 *   node_age [dev] N int64 seconds ((now - creation).seconds, scaler.py:78),
 *   any_pending = len(pods_to_schedule) > 0, idle_threshold [host] S int64,
 *   out_state [dev] S x N bytes.
 */
ACSFIT_API acsfit_status acsfit_node_states(acsfit_ctx *ctx, const int64_t *row_ptr, const int32_t *run_idx,
                                 const double *req_run, const uint8_t *flags_run,
                                 const double *cap_type, const int32_t *node_type,
                                 const uint8_t *node_flags, const int64_t *node_age, int64_t N,
                                 int D, int any_pending, const int64_t *idle_threshold, int S,
                                 uint8_t *out_state, acsfit_stream_t stream);

/*
 * K6b maintain_actions -- the decision part of EngineScaler.maintain (reference
 * engine_scaler.py:133-182): SPARE_AGENT substitution against the per-pool drain budget
 * (budget0 = actual - len(unschedulable) - spare, :136; it only decrements when not dry_run,
 * :154-159) and the state -> action table.
 *   io_state [dev] N bytes (rewritten; nodes of non-scalable pools become ACSFIT_ST_NOT_EVALUATED)
 *   node_pool [dev] N int32, budget0 [host] T int64, pool_scalable [host] T bytes,
 *   out_action [dev] N bytes (ACSFIT_ACT_*).
 */
ACSFIT_API acsfit_status acsfit_maintain_actions(acsfit_ctx *ctx, uint8_t *io_state, const int32_t *node_pool,
                                      int64_t N, const int64_t *budget0,
                                      const uint8_t *pool_scalable, int T, int dry_run,
                                      uint8_t *out_action, acsfit_stream_t stream);

/*
 * Fused scale-up tick on DEVICE buffers: get_pods_to_schedule + get_pending_pods +
 * fulfill_pending (reference cluster.py:169-175, :206-215).  Precondition: pod uids are unique
 * (otherwise the host layer applies the dict collapse of scaler.py:119 between the individual
 * entry points).  D must be 2, 4, 8 or 16.
 *   in : req [dev] P x D (pending-unassigned pods, list order)
 *        unit_all [host] T x D in agent_pools order (feasibility, capacity.py:24-32)
 *        unit_ordered / pool_actual / pool_max / pool_ignored [host] in VISITING order
 *        cap_type [dev] K x D, node_type [dev] N, used_inout [dev] N x D
 *   out: out_feasible [dev] P bytes
 *        out_placed   [dev] P int32: node index, -1 = pending, -2 = infeasible (skipped)
 *        out_new_size / out_units_needed / out_bins_opened [host] T int64
 *        out_acc_pool [dev] P int32: visiting index of the pool that accounted the pod, else -1
 *        out_counters [host] 4 x uint64 = {pods to schedule, pending, num_unaccounted, decisions}
 *        where decisions = every `.possible` / can_fit evaluation the reference would perform.
 */
ACSFIT_API acsfit_status acsfit_scale_up(acsfit_ctx *ctx, const double *req, int64_t P, int D,
                              const double *unit_all, const double *unit_ordered,
                              const int32_t *pool_actual, const int32_t *pool_max,
                              const uint8_t *pool_ignored, int T, int64_t over_provision,
                              const double *cap_type, const int32_t *node_type, double *used_inout,
                              int64_t N, uint8_t *out_feasible, int32_t *out_placed,
                              int64_t *out_new_size, int64_t *out_units_needed,
                              int64_t *out_bins_opened, int32_t *out_acc_pool,
                              uint64_t *out_counters, acsfit_stream_t stream);

/*
 * Host-buffer entry points (the "plugin" calls): the same computations with HOST pointers
 * everywhere; the library copies H2D (fastest from pinned memory), runs the kernels on the
 * default stream and copies the results back before returning.
 * acsfit_scale_up_host = acsfit_scale_up with host buffers (cap_type has K rows).
 */
ACSFIT_API acsfit_status acsfit_scale_up_host(acsfit_ctx *ctx, const double *req, int64_t P, int D,
                                   const double *unit_all, const double *unit_ordered,
                                   const int32_t *pool_actual, const int32_t *pool_max,
                                   const uint8_t *pool_ignored, int T, int64_t over_provision,
                                   const double *cap_type, int K, const int32_t *node_type,
                                   double *used_inout, int64_t N, uint8_t *out_feasible,
                                   int32_t *out_placed, int64_t *out_new_size,
                                   int64_t *out_units_needed, int64_t *out_bins_opened,
                                   int32_t *out_acc_pool, uint64_t *out_counters);

/* acsfit_occupancy_host = acsfit_occupancy with host buffers (reference cluster.py:165-168): R rows of req_run,
 * used_inout N x D is read, updated and written back. */
ACSFIT_API acsfit_status acsfit_occupancy_host(acsfit_ctx *ctx, const int64_t *row_ptr, const int32_t *run_idx,
                                    const double *req_run, int64_t R, int64_t N, int D, double *used_inout);

/* acsfit_maintain_host = node_states (one threshold) + maintain_actions with host buffers
 * (reference engine_scaler.py:120-182, scaler.py:61-114). out_state/out_action N bytes. */
ACSFIT_API acsfit_status acsfit_maintain_host(acsfit_ctx *ctx, const int64_t *row_ptr, const int32_t *run_idx,
                                   const double *req_run, const uint8_t *flags_run, int64_t R,
                                   const double *cap_type, int K, const int32_t *node_type,
                                   const uint8_t *node_flags, const int64_t *node_age,
                                   const int32_t *node_pool, int64_t N, int D, int any_pending,
                                   int64_t idle_threshold, const int64_t *budget0,
                                   const uint8_t *pool_scalable, int T, int dry_run,
                                   uint8_t *out_state, uint8_t *out_action);

/*
 * Cluster mode: ONE cluster (one pod list, one node list) on the `world` GPUs of one box, one process and one
 * ctx per GPU.  The two sequential first-fit loops (reference cluster.py:184-204 and scaler.py:131-147) are
 * parallel over the NODE (bin) axis, so rank r owns the r-th contiguous range of nodes (of every bin pass) and
 * the pipeline of stages simply continues on the next GPU: stage 0 of rank r polls the progress counter of the
 * last stage of rank r-1 and reads that rank's "still unplaced" bitmap, both through NVLink peer memory
 * (CUDA IPC), tile by tile, inside the kernel.  Pods visit nodes in list order and every node sees the pods in
 * list order, so placements, `used` bit patterns, pool sizes and the credited decisions are exactly those of
 * the single-GPU call.  The only other exchanges are in-stream peer-memory barriers and element-wise merges of
 * the placement vectors (each pod is placed by exactly one rank).
 *
 * Protocol: every rank calls acsfit_cluster_init, the 64-byte handles are all-gathered by the host layer
 * (torch.distributed) and passed to acsfit_cluster_connect.  From then on acsfit_first_fit_nodes,
 * acsfit_fulfill_pending, acsfit_scale_up and acsfit_scale_up_host must be called by ALL ranks with IDENTICAL
 * arguments (replicated inputs); every rank returns the complete, identical result.
 *   max_pods / max_nodes / max_dims size the exchange region this rank exports (cudaMalloc'ed by the library).
 */
#define ACSFIT_IPC_HANDLE_BYTES 64
#define ACSFIT_MAX_RANKS 8
ACSFIT_API acsfit_status acsfit_cluster_init(acsfit_ctx *ctx, int rank, int world, int64_t max_pods,
                                             int64_t max_nodes, int max_dims, void *out_handle);
/* all_handles [host]: world x ACSFIT_IPC_HANDLE_BYTES, rank-major (this rank's own entry included) */
ACSFIT_API acsfit_status acsfit_cluster_connect(acsfit_ctx *ctx, const void *all_handles);
/* in-stream barrier over the ranks (peer-memory flags); the first call must follow a host-level barrier */
ACSFIT_API acsfit_status acsfit_cluster_barrier(acsfit_ctx *ctx, acsfit_stream_t stream);
/* geometry of the last cluster-mode node pass: stage width, stages of this rank, pod blocks, resident stage CTAs */
ACSFIT_API acsfit_status acsfit_cluster_last_plan(const acsfit_ctx *ctx, int *out_tn, int *out_stages,
                                                  int *out_blocks, int *out_resident);

/* number of kernels this ctx has launched so far (bench.py reports it as gpu_launches) */
ACSFIT_API uint64_t acsfit_launch_count(const acsfit_ctx *ctx);
/* CUDA-event time (ms) of the last first-fit / bin-pack pipeline kernel sequence and the
 * decisions it evaluated, measured on the launching stream (bench.py roofline leg) */
ACSFIT_API acsfit_status acsfit_last_pipeline_stats(const acsfit_ctx *ctx, double *out_ms,
                                         uint64_t *out_decisions, int *out_stages, int *out_tiles);

/* developer probes (tools/perf_probe.py): per-stage clock64 phase totals [stages][8] of the last
 * pipeline launch (wait, load, scan, resolve, publish, refresh, hits, tiles) and a per-tile trace
 * [tiles][8] of one chosen stage.  Enabling them adds a clock read per phase on thread 0. */
ACSFIT_API acsfit_status acsfit_debug_profile(acsfit_ctx *ctx, int enabled, unsigned long long *out,
                                              int max_stages, int *out_stages);
ACSFIT_API acsfit_status acsfit_debug_trace(acsfit_ctx *ctx, int stage, unsigned long long *out,
                                            int max_tiles);

#ifdef __cplusplus
}
#endif
#endif /* ACSFIT_H */
