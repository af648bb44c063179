"""Snapshot flattening: kube objects <-> dense float64 device buffers (host logic, no arithmetic).

One tick's cluster state is turned into contiguous arrays once, shipped to HBM, and the
decisions come back as index vectors that are mapped onto the original KubePod / KubeNode
objects, including `node.used_capacity`, so downstream Python sees the state the reference
would have produced (SURVEY.md section 3.2 "device boundary").

Layout: resource dims = sorted union of keys over every vector involved; a missing key is 0.0
(bit-exact with KubeResource's own missing-key rule, kube.py:204-206,:210-212); the column count
is padded with zero columns to 2/4/8/16 for the first-fit kernels.
"""
import numpy as np
import torch

from .engine import Engine, padded_dims
from .kube import _hostfast

_engine = None


def get_engine():
    """the process-wide CUDA engine; raises when libacsfit.so or a GPU is missing (no fallback)."""
    global _engine
    if _engine is None:
        _engine = Engine()
    return _engine


def set_engine(engine):
    """install a specific engine (e.g. one bound to this rank's GPU)."""
    global _engine
    _engine = engine
    return engine


def _group(objects):
    """(group index per object, one representative per group): identity grouping in first-occurrence order."""
    objects = objects if isinstance(objects, list) else list(objects)
    inv = np.empty(len(objects), dtype=np.int64)
    if _hostfast is not None:
        return inv, _hostfast.group_ids(objects, inv)
    slot, uniq = {}, []
    for i, obj in enumerate(objects):
        j = slot.get(id(obj))
        if j is None:
            j = slot[id(obj)] = len(uniq)
            uniq.append(obj)
        inv[i] = j
    return inv, uniq


class Dims(object):
    """column order of one flattening.  Resource vectors are deduplicated by object identity first: pods of one
    template share their KubeResource (kube._pod_resources), so 10^5 pods flatten as a few dozen distinct rows
    plus one fancy-index, not as 10^5 Python dict walks."""

    def __init__(self, resources):
        keys = set()
        for res in _group(resources)[1]:
            keys.update(res.raw.keys())
        self.keys = sorted(keys)
        self.index = {k: i for i, k in enumerate(self.keys)}
        self.D = max(1, len(self.keys))
        self.Dp = padded_dims(self.D)

    def rows(self, resources, n=None, groups=False):
        """[len(resources), Dp] float64 rows; with groups=True also (group index per row, the distinct vectors)."""
        resources = list(resources)
        index = self.index
        inv, uniq = _group(resources)
        urows = np.zeros((len(uniq), self.Dp), dtype=np.float64)
        for j, res in enumerate(uniq):
            for key, value in res.raw.items():
                urows[j, index[key]] = value
        out = np.zeros((len(resources) if n is None else n, self.Dp), dtype=np.float64)
        if len(resources):
            out[:len(resources)] = urows[inv]
        return (out, inv, uniq) if groups else out


def _touched_keys(node_of, inv, uniq):
    """{node: union of the key sets of the vectors counted on it}: one set update per distinct (node, vector) pair
    instead of one per pod (the pods of a template share their vector)."""
    touched = {}
    if not len(node_of):
        return touched
    G = max(1, len(uniq))
    pairs = np.unique(np.asarray(node_of, dtype=np.int64) * G + np.asarray(inv, dtype=np.int64))
    keysets = [set(u.raw) for u in uniq]
    for pair in pairs.tolist():
        n, g = divmod(pair, G)
        have = touched.get(n)
        if have is None:
            touched[n] = set(keysets[g])
        else:
            have.update(keysets[g])
    return touched


def _check_finite(a, what):
    if not np.isfinite(a).all():
        raise ValueError("non-finite %s is outside the supported domain" % what)


def _nonfinite_rows(req):
    return ~np.isfinite(req).all(axis=1)


def _node_types(nodes, dims):
    """(cap_type[K, Dp], node_type[N]) -- one capacity row per distinct KubeResource object/type."""
    rows, index, node_type = [], {}, np.zeros(len(nodes), dtype=np.int32)
    for i, node in enumerate(nodes):
        key = id(node.capacity)
        if key not in index:
            index[key] = len(rows)
            rows.append(node.capacity)
        node_type[i] = index[key]
    cap = dims.rows(rows)
    _check_finite(cap, "node capacity")
    return cap, node_type


def _write_back_used(nodes, dims, used, touched_keys):
    """store used[n] into node.used_capacity with the key set the reference would hold: the
    node's previous keys plus the keys of every pod counted on it (KubeResource.__add__ unions)."""
    from .kube import KubeResource
    for n, keys in touched_keys.items():
        node = nodes[n]
        merged = set(node.used_capacity.raw) | keys
        res = KubeResource()
        res.raw = {k: float(used[n, dims.index[k]]) for k in merged}
        node.used_capacity = res


# ------------------------------------------------------------------------------------------------
# K0: feasibility (capacity.is_possible over the pools)
# ------------------------------------------------------------------------------------------------
def feasible_pods(pods, agent_pools):
    """bool[len(pods)]: (RESOURCE_SPEC[pool.instance_type] - pod.resources).possible for some pool."""
    from . import capacity
    pods = list(pods)
    if not pods:
        return np.zeros(0, dtype=bool)
    units = [capacity.RESOURCE_SPEC[pool.instance_type] for pool in agent_pools]
    if not units:
        return np.zeros(len(pods), dtype=bool)
    dims = Dims([p.resources for p in pods] + units)
    req = dims.rows(p.resources for p in pods)
    unit = dims.rows(units)
    _check_finite(unit, "unit capacity")
    bad = _nonfinite_rows(req)       # an infinite request fits nowhere (cap - inf < 0)
    req[bad] = 0.0
    eng = get_engine()
    mask, _ = eng.feasible_mask(eng.dev(req, torch.float64), eng.dev(unit, torch.float64))
    out = mask.cpu().numpy().astype(bool)
    out[bad] = False
    return out


# ------------------------------------------------------------------------------------------------
# K1: occupancy (node.count_pod over the running pods)
# ------------------------------------------------------------------------------------------------
def running_csr(nodes, running_pods):
    """CSR of running pods per node in pod-list order (cluster.py:165-168 / engine_scaler.py:129-131).
    Two nodes with the same name both get the pod, as the reference's name comparison does."""
    by_name = {}
    for j, pod in enumerate(running_pods):
        by_name.setdefault(pod.node_name, []).append(j)
    row_ptr = np.zeros(len(nodes) + 1, dtype=np.int64)
    idx = []
    for n, node in enumerate(nodes):
        lst = by_name.get(node.name, ())
        idx.extend(lst)
        row_ptr[n + 1] = len(idx)
    return row_ptr, np.asarray(idx, dtype=np.int32)


def count_running_pods(nodes, running_pods):
    """the occupancy loop of Cluster.loop_logic on the GPU; updates node.used_capacity."""
    nodes, running_pods = list(nodes), list(running_pods)
    if not nodes or not running_pods:
        return
    row_ptr, run_idx = running_csr(nodes, running_pods)
    if not len(run_idx):
        return
    dims = Dims([p.resources for p in running_pods] + [n.used_capacity for n in nodes])
    req_run, inv_run, uniq_run = dims.rows((running_pods[j].resources for j in run_idx), groups=True)  # rows in node order: a contiguous table
    if (req_run < 0).any() or np.isnan(req_run).any():
        raise ValueError("negative or NaN resource request")
    used = dims.rows(n.used_capacity for n in nodes)
    eng = get_engine()
    d_used = eng.dev(used, torch.float64)
    eng.occupancy(eng.dev(row_ptr, torch.int64), None, eng.dev(req_run, torch.float64), d_used)
    used = d_used.cpu().numpy()
    node_of_row = np.repeat(np.arange(len(nodes), dtype=np.int64), np.diff(row_ptr))
    _write_back_used(nodes, dims, used, _touched_keys(node_of_row, inv_run, uniq_run))


# ------------------------------------------------------------------------------------------------
# first fit over nodes (Cluster.get_pending_pods)
# ------------------------------------------------------------------------------------------------
def first_fit_nodes(pods, nodes):
    """returns placed[len(pods)] (node index or -1) and mutates node.used_capacity like count_pod."""
    pods, nodes = list(pods), list(nodes)
    placed = np.full(len(pods), -1, dtype=np.int32)
    if not pods or not nodes:
        return placed
    dims = Dims([p.resources for p in pods] + [n.capacity for n in nodes] + [n.used_capacity for n in nodes])
    req, inv_req, uniq_req = dims.rows((p.resources for p in pods), groups=True)
    bad = _nonfinite_rows(req)  # infinite request: can_fit is False on every node
    ok_idx = np.nonzero(~bad)[0].astype(np.int32)
    cap, node_type = _node_types(nodes, dims)
    used = dims.rows(n.used_capacity for n in nodes)
    req[bad] = 0.0
    eng = get_engine()
    d_used = eng.dev(used, torch.float64)
    d_placed, _ = eng.first_fit_nodes(eng.dev(req, torch.float64),
                                      None if not bad.any() else eng.dev(ok_idx, torch.int32),
                                      eng.dev(cap, torch.float64), eng.dev(node_type, torch.int32), d_used)
    res = d_placed.cpu().numpy()
    if bad.any():
        placed[ok_idx] = res
    else:
        placed = res
    used = d_used.cpu().numpy()
    sel = np.nonzero(placed >= 0)[0]
    _write_back_used(nodes, dims, used, _touched_keys(placed[sel], inv_req[sel], uniq_req))
    return placed


def single_can_fit(node, resources):
    """KubeNode.can_fit for one vector: a 1 x 1 first fit on a scratch copy of the node row."""
    from .kube import KubeResource

    class _P(object):
        pass
    pod = _P()
    pod.resources = resources

    class _N(object):
        pass
    scratch = _N()
    scratch.capacity = node.capacity
    scratch.used_capacity = KubeResource()
    scratch.used_capacity.raw = dict(node.used_capacity.raw)
    return bool(first_fit_nodes([pod], [scratch])[0] >= 0)


# ------------------------------------------------------------------------------------------------
# fulfill_pending (bins)
# ------------------------------------------------------------------------------------------------
def fulfill(unique_pods, num_listed, ordered_pools, ignored_pool_names, over_provision):
    """Scaler.fulfill_pending's packing + size arithmetic; pools already in visiting order."""
    from . import capacity
    units = [capacity.RESOURCE_SPEC[pool.instance_type] for pool in ordered_pools]
    dims = Dims([p.resources for p in unique_pods] + units)
    req = dims.rows(p.resources for p in unique_pods)
    unit = dims.rows(units) if units else np.zeros((0, dims.Dp))
    _check_finite(unit, "unit capacity")
    _check_finite(req, "pending pod request")  # pending pods passed is_possible, so they are finite
    eng = get_engine()
    res = eng.fulfill_pending(
        eng.dev(req, torch.float64), num_listed, unit,
        np.asarray([pool.actual_capacity for pool in ordered_pools], dtype=np.int32),
        np.asarray([min(pool.max_size, 2 ** 31 - 1) for pool in ordered_pools], dtype=np.int32),
        np.asarray([1 if pool.name in ignored_pool_names else 0 for pool in ordered_pools], dtype=np.uint8),
        over_provision)
    res["acc_pool"] = res["acc_pool"].cpu().numpy()
    res["bin_of"] = res["bin_of"].cpu().numpy()
    return res


# ------------------------------------------------------------------------------------------------
# idle scan + maintain decisions
# ------------------------------------------------------------------------------------------------
POD_BUSY, POD_UNDRAINABLE = 1, 2
NODE_UNSCHEDULABLE = 1


def pod_flags(pod, now_by_tz=None):
    """host-computed booleans of get_node_state (scaler.py:76, :82-83).  `now_by_tz` (a dict) lets a batch read
    the clock once per time zone instead of once per pod (kube.py:68 reads it per call; within one maintain pass
    the readings differ by microseconds against a one-hour grace period)."""
    proxy = 'kube-proxy' in pod.name
    flags = 0
    if not pod.is_mirrored() and not proxy:
        flags |= POD_BUSY
    if proxy:
        return flags
    if now_by_tz is None:
        drainable = pod.is_drainable()
    else:  # is_drainable(), with the clock of the batch (same short-circuit order: replicated, critical, grace)
        drainable = pod.is_replicated() and not pod.is_critical()
        if drainable:
            start = pod.start_time
            if not start:
                drainable = False
            else:
                tz = start.tzinfo
                now = now_by_tz.get(id(tz))  # (tzinfo objects need not be hashable: keyed by identity)
                if now is None:
                    from . import utils
                    now = now_by_tz[id(tz)] = utils.now(tz)
                drainable = not ((now - start) < pod._DRAIN_GRACE_PERIOD)
    if not drainable:
        flags |= POD_UNDRAINABLE
    return flags


def node_age_seconds(node, now_by_tz=None):
    """(now - creation).seconds -- wraps at one day, like the reference (scaler.py:78).  `now_by_tz`: one clock
    reading per time zone and batch, as in pod_flags (now(tz) on a dateutil zone costs ~10 us: it converts from UTC)."""
    from . import utils
    tz = node.creation_time.tzinfo
    if now_by_tz is None:
        return (utils.now(tz) - node.creation_time).seconds
    now = now_by_tz.get(id(tz))
    if now is None:
        now = now_by_tz[id(tz)] = utils.now(tz)
    return (now - node.creation_time).seconds


def node_states(nodes, pods_lists, any_pending, idle_threshold):
    """state code per node (include/acsfit.h ACSFIT_ST_*); pods_lists[n] = the node's pods in order."""
    nodes = list(nodes)
    if not nodes:
        return np.zeros(0, dtype=np.uint8)
    flat, row_ptr = [], np.zeros(len(nodes) + 1, dtype=np.int64)
    for n, lst in enumerate(pods_lists):
        flat.extend(lst)
        row_ptr[n + 1] = len(flat)
    dims = Dims([p.resources for p in flat] + [n.capacity for n in nodes])
    req_run = dims.rows(p.resources for p in flat) if flat else np.zeros((0, dims.Dp))
    if (req_run < 0).any() or np.isnan(req_run).any():
        raise ValueError("negative or NaN resource request")
    clock = {}
    flags = np.asarray([pod_flags(p, clock) for p in flat], dtype=np.uint8)
    cap, node_type = _node_types(nodes, dims)
    node_flags = np.asarray([NODE_UNSCHEDULABLE if n.unschedulable else 0 for n in nodes], dtype=np.uint8)
    age = np.asarray([node_age_seconds(n, clock) for n in nodes], dtype=np.int64)
    eng = get_engine()
    st = eng.node_states(eng.dev(row_ptr, torch.int64),
                         None,  # contiguous table: the rows ARE in node order (bulk-copy streaming kernel)
                         eng.dev(req_run, torch.float64), eng.dev(flags, torch.uint8),
                         eng.dev(cap, torch.float64), eng.dev(node_type, torch.int32),
                         eng.dev(node_flags, torch.uint8), eng.dev(age, torch.int64), any_pending,
                         np.asarray([idle_threshold], dtype=np.int64))
    return st[0].contiguous()


def maintain_actions(state_dev, node_pool, budget0, pool_scalable, dry_run):
    eng = get_engine()
    state, action = eng.maintain_actions(state_dev, eng.dev(np.asarray(node_pool, dtype=np.int32), torch.int32),
                                         np.asarray(budget0, dtype=np.int64),
                                         np.asarray(pool_scalable, dtype=np.uint8), dry_run)
    return state.cpu().numpy(), action.cpu().numpy()
