"""Seeded synthetic cluster snapshots on the dense layout.

Used by bench.py and the parity tests.  Values are formed exactly the way the reference's
parser forms them (utils.py:36-42: float(digits) * multiplier), e.g. 100m -> 100.0 * 1e-3,
64Mi -> 64.0 * 2**20, so the float64 bit patterns are the ones a real cluster would produce.

Dimension order (sorted key union, as the host layer flattens): D=4 ->
[alpha.kubernetes.io/nvidia-gpu, cpu, memory, pods]; D=8 appends four synthetic extended
resources x/res0..x/res3.
"""
import datetime
import json

import numpy as np

GPU, CPU, MEM, PODS = 0, 1, 2, 3
DIM_NAMES_4 = ["alpha.kubernetes.io/nvidia-gpu", "cpu", "memory", "pods"]
DIM_NAMES_8 = DIM_NAMES_4 + ["x/res0", "x/res1", "x/res2", "x/res3"]

# (name, gpu, cpu, memory, pods) -- values of the reference's data/capacity.json for these types,
# listed in that file's order (= cost order, capacity.py:34-36)
INSTANCE_TYPES = [
    ("Standard_D2_v2", 0.0, 2.0, 7096762368.0, 110.0),
    ("Standard_D4_v3", 0.0, 4.0, 16760438784.0, 110.0),
    ("Standard_D8s_v3", 0.0, 8.0, 33940307968.0, 110.0),
    ("Standard_D16s_v3", 0.0, 16.0, 68300046336.0, 110.0),
    ("Standard_NC6", 1.0, 6.0, 59087724544.0, 110.0),
    ("Standard_NC12", 2.0, 12.0, 114688000000.0, 110.0),
    ("Standard_E32_v3", 0.0, 32.0, 274458476544.0, 110.0),
    ("Standard_NC24", 4.0, 24.0, 229376000000.0, 110.0),
]

CPU_MILLI = np.array([100, 250, 500, 1000, 1500, 2000, 4000], dtype=np.float64)
MEM_MIB = np.array([64, 128, 256, 512, 1024, 4096], dtype=np.float64)


def capacity_rows(T, D):
    """cap_type [T, D]: one instance type per pool, pool t uses INSTANCE_TYPES[t]."""
    assert 1 <= T <= len(INSTANCE_TYPES) and D in (4, 8)
    cap = np.zeros((T, D), dtype=np.float64)
    for t in range(T):
        _, gpu, cpu, mem, pods = INSTANCE_TYPES[t]
        cap[t, GPU], cap[t, CPU], cap[t, MEM], cap[t, PODS] = gpu, cpu, mem, pods
        if D == 8:
            cap[t, 4:8] = np.array([8.0, 16.0, 4.0, 64.0]) * (1 + t)
    return cap


def pod_rows(rng, n, D, gpu_prob=0.1):
    """request rows as KubePod.resources would hold them (kube.py:41-49: pods=1 + sum of requests)."""
    req = np.zeros((n, D), dtype=np.float64)
    req[:, CPU] = CPU_MILLI[rng.integers(0, len(CPU_MILLI), size=n)] * 1e-3
    req[:, MEM] = MEM_MIB[rng.integers(0, len(MEM_MIB), size=n)] * float(2 ** 20)
    req[:, GPU] = (rng.random(n) < gpu_prob).astype(np.float64)
    req[:, PODS] = 1.0
    if D == 8:
        extra = rng.integers(1, 5, size=(n, 4)).astype(np.float64)
        extra[rng.random((n, 4)) < 0.7] = 0.0
        req[:, 4:8] = extra
    return req


def make_cluster(P, N, D=4, T=1, seed=0, free_frac=0.15, run_per_node=14, gpu_prob=0.1,
                 ignored=(), max_size=None, over_provision=0):
    """a full tick snapshot.

    free_frac: fraction of nodes that are only partly filled (the rest are filled by first-fit
    until the next running pod no longer fits), which sets how many pending pods find a node.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    cap_type = capacity_rows(T, D)
    node_type = (np.arange(N, dtype=np.int64) % T).astype(np.int32)  # types cycled over the pools
    node_pool = node_type.copy()

    # running pods: per node up to run_per_node candidates, keep the prefix that fits (ordered sums)
    cand = pod_rows(rng, N * run_per_node, D, gpu_prob=0.02).reshape(N, run_per_node, D)
    # bigger instance types run proportionally bigger pods (integer multiples, so the values are
    # still float(n) * multiplier), otherwise 14 pods could never fill a 32-core node
    scale = np.maximum(1.0, np.floor(cap_type[node_type, CPU] / 2.0))
    cand[:, :, CPU] = (cand[:, :, CPU] * 1e3).round() * scale[:, None] * 1e-3
    cand[:, :, MEM] = cand[:, :, MEM] * scale[:, None]
    csum = np.zeros((N, D), dtype=np.float64)
    keep = np.zeros((N, run_per_node), dtype=bool)
    ok = np.ones(N, dtype=bool)
    partial = rng.random(N) < free_frac
    limit = np.where(partial, rng.integers(0, max(2, run_per_node // 3), size=N), run_per_node)
    cap_n = cap_type[node_type]
    for k in range(run_per_node):
        nxt = csum + cand[:, k, :]
        fits = np.all(cap_n - nxt >= 0, axis=1)
        ok = ok & fits & (k < limit)
        keep[:, k] = ok
        csum = np.where(ok[:, None], nxt, csum)
    counts = keep.sum(axis=1).astype(np.int64)
    row_ptr = np.zeros(N + 1, dtype=np.int64)
    np.cumsum(counts, out=row_ptr[1:])
    req_run = cand[keep]  # row-major over (node, k): grouped by node, list order inside a node
    R = req_run.shape[0]
    run_idx = np.arange(R, dtype=np.int32)
    # pod flags: 5% DaemonSet (not busy), 10% bare pods (undrainable)
    flags_run = np.full(R, 1, dtype=np.uint8)
    r = rng.random(R)
    flags_run[r < 0.05] = 0
    flags_run[(r >= 0.05) & (r < 0.15)] |= 2
    node_flags = (rng.random(N) < 0.02).astype(np.uint8)
    # ages log-uniform 1 s .. 3 d, then timedelta.seconds wraps at one day (scaler.py:78)
    age = np.exp(rng.uniform(0.0, np.log(3 * 86400.0), size=N)).astype(np.int64) % 86400

    req = pod_rows(rng, P, D, gpu_prob=gpu_prob)
    pool_actual = np.bincount(node_pool, minlength=T).astype(np.int32)
    pool_max = np.full(T, int(max_size if max_size is not None else P + N + 1000), dtype=np.int32)
    pool_ignored = np.zeros(T, dtype=np.uint8)
    for t in ignored:
        pool_ignored[t] = 1
    return {
        "P": P, "N": N, "D": D, "T": T, "seed": seed,
        "dim_names": DIM_NAMES_4 if D == 4 else DIM_NAMES_8,
        "req": req, "cap_type": cap_type, "node_type": node_type, "node_pool": node_pool,
        "row_ptr": row_ptr, "run_idx": run_idx, "req_run": req_run, "flags_run": flags_run,
        "node_flags": node_flags, "node_age": age,
        # pools: agent_pools order == visiting order here (types are listed in cost order)
        "unit_all": cap_type.copy(), "unit_ordered": cap_type.copy(),
        "pool_actual": pool_actual, "pool_max": pool_max, "pool_ignored": pool_ignored,
        "over_provision": int(over_provision),
    }


def make_idle_cluster(N, D=4, T=1, seed=0, mean_pods=10.0):
    """BASELINE config 5: N nodes with a running-pod occupancy list of ~Poisson(mean_pods) small pods each
    (clipped to the 110-pod limit), 5 % DaemonSet pods, 10 % bare (undrainable) pods, ages log-uniform
    1 s .. 3 d (wrapped like timedelta.seconds), 2 % cordoned nodes."""
    rng = np.random.Generator(np.random.PCG64(seed))
    cap_type = capacity_rows(T, D)
    node_type = (np.arange(N, dtype=np.int64) % T).astype(np.int32)
    counts = np.minimum(rng.poisson(mean_pods, size=N), 110).astype(np.int64)
    row_ptr = np.zeros(N + 1, dtype=np.int64)
    np.cumsum(counts, out=row_ptr[1:])
    R = int(row_ptr[-1])
    req_run = np.zeros((R, D), dtype=np.float64)
    scale = np.repeat(np.maximum(1.0, np.floor(cap_type[node_type, CPU] / 2.0)), counts)
    req_run[:, CPU] = np.array([20, 50, 100, 150], dtype=np.float64)[rng.integers(0, 4, size=R)] * scale * 1e-3
    req_run[:, MEM] = np.array([32, 64, 128, 256], dtype=np.float64)[rng.integers(0, 4, size=R)] * float(2 ** 20) * scale
    req_run[:, PODS] = 1.0
    if D == 8:
        extra = rng.integers(1, 3, size=(R, 4)).astype(np.float64)
        extra[rng.random((R, 4)) < 0.8] = 0.0
        req_run[:, 4:8] = extra
    flags_run = np.full(R, 1, dtype=np.uint8)
    r = rng.random(R)
    flags_run[r < 0.05] = 0
    flags_run[(r >= 0.05) & (r < 0.15)] |= 2
    return {
        "N": N, "D": D, "T": T, "P": 0, "seed": seed, "cap_type": cap_type, "node_type": node_type,
        "node_pool": node_type.copy(), "row_ptr": row_ptr, "run_idx": np.arange(R, dtype=np.int32),
        "req_run": req_run, "flags_run": flags_run,
        "node_flags": (rng.random(N) < 0.02).astype(np.uint8),
        "node_age": np.exp(rng.uniform(0.0, np.log(3 * 86400.0), size=N)).astype(np.int64) % 86400,
        "pool_actual": np.bincount(node_type, minlength=T).astype(np.int32),
    }


def initial_used(c):
    """used[N, D] after the occupancy loop (cluster.py:165-168), computed with ordered numpy sums
    (host helper for building inputs; the device path is acsfit_occupancy)."""
    N, D = c["N"], c["D"]
    used = np.zeros((N, D), dtype=np.float64)
    counts = np.diff(c["row_ptr"])
    maxc = int(counts.max()) if N else 0
    start = c["row_ptr"][:-1]
    for k in range(maxc):
        sel = counts > k
        rows = c["req_run"][c["run_idx"][start[sel] + k]]
        used[sel] = used[sel] + rows
    return used


# ------------------------------------------------------------------------------------------------
# the same snapshot as kube-API style objects (what pykube would list), for the Python entry surface
# ------------------------------------------------------------------------------------------------
NOW = datetime.datetime(2026, 9, 21, 12, 0, 0, tzinfo=datetime.timezone.utc)
CLUSTER_ID = "16334397"


def quantity(dim_name, value):
    """the kube quantity string that utils.parse_SI turns back into `value` (checked)."""
    if dim_name == "cpu":
        n = int(round(value * 1000.0))
        s, back = "%dm" % n, float(n) * 1e-3
    elif dim_name == "memory":
        n = int(round(value / float(2 ** 20)))
        s, back = "%dMi" % n, float(n) * float(2 ** 20)
    else:
        n = int(round(value))
        s, back = "%d" % n, float(n)
    if back != value:
        raise ValueError("%r of %s has no exact quantity string" % (value, dim_name))
    return s


def requests_of(row, dim_names):
    """container `requests` dict of one dense row (the `pods` column is KubePod's own pods=1, kube.py:49)."""
    return {name: quantity(name, float(v)) for name, v in zip(dim_names, row) if name != "pods" and v != 0.0}


def iso(dt):
    return dt.strftime("%Y-%m-%dT%H:%M:%SZ")


def pool_name(t):
    return "pool%d" % t


def node_name(t, i):
    return "k8s-%s-%s-%d" % (pool_name(t), CLUSTER_ID, i)


def kube_objects(c):
    """synthetic cluster dict -> {'nodes': [...], 'pods': [...], 'arm_parameters': {...}} of kube-style dicts.
    Pod list order: the running pods in CSR order (per node, list order), then the pending pods."""
    names = c["dim_names"]
    N, T = c["N"], c["T"]
    nodes = []
    for i in range(N):
        t = int(c["node_type"][i])
        created = NOW - datetime.timedelta(seconds=int(c["node_age"][i]))
        spec = {"unschedulable": True} if c["node_flags"][i] & 1 else {}
        nodes.append({"metadata": {"name": node_name(t, i), "creationTimestamp": iso(created),
                                   "labels": {"beta.kubernetes.io/instance-type": INSTANCE_TYPES[t][0],
                                              "failure-domain.beta.kubernetes.io/region": "southcentralus",
                                              "kubernetes.io/hostname": node_name(t, i)}},
                      "spec": spec})
    pods = []
    old = iso(NOW - datetime.timedelta(hours=5))
    row_ptr, run_idx = c["row_ptr"], c["run_idx"]
    for i in range(N):
        t = int(c["node_type"][i])
        for k in range(int(row_ptr[i]), int(row_ptr[i + 1])):
            j = int(run_idx[k])
            f = int(c["flags_run"][j])
            md = {"name": "run-%d" % j, "namespace": "default", "uid": "run-%d" % j, "creationTimestamp": old}
            # flag bits (include/acsfit.h): 1 = busy (not mirrored), 2 = undrainable; kube.py:51-71
            if f == 0:    # DaemonSet pod: mirrored (not busy) and replicated (drainable)
                md["annotations"] = {"kubernetes.io/created-by": json.dumps(
                    {"kind": "SerializedReference", "reference": {"kind": "DaemonSet", "name": "ds"}})}
            elif f == 2:  # static (mirror) pod: not busy, but not drainable either
                md["annotations"] = {"kubernetes.io/config.mirror": "x"}
            elif f == 1:  # ReplicaSet pod: busy and drainable
                md["annotations"] = {"kubernetes.io/created-by": json.dumps(
                    {"kind": "SerializedReference", "reference": {"kind": "ReplicaSet", "name": "rs"}})}
            # f == 3: a bare pod, busy and undrainable -- no annotation
            pods.append({"metadata": md,
                         "spec": {"nodeName": node_name(t, i),
                                  "containers": [{"name": "c", "resources": {"requests": requests_of(c["req_run"][j], names)}}]},
                         "status": {"phase": "Running", "startTime": old}})
    for p in range(c["P"]):
        pods.append({"metadata": {"name": "pend-%d" % p, "namespace": "default", "uid": "pend-%d" % p,
                                  "creationTimestamp": old},
                     "spec": {"containers": [{"name": "c", "resources": {"requests": requests_of(c["req"][p], names)}}]},
                     "status": {"phase": "Pending"}})
    arm = {"masterVMSize": {"value": "Standard_D2_v2"}}
    for t in range(T):
        arm[pool_name(t) + "Count"] = {"value": 1}
        arm[pool_name(t) + "VMSize"] = {"value": INSTANCE_TYPES[t][0]}
    return {"nodes": nodes, "pods": pods, "arm_parameters": arm}


