#!/usr/bin/env python3
"""TEST INFRASTRUCTURE -- generate tests/golden/*.json by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python oracle/make_golden.py

It imports `/root/reference/autoscaler` under the sys.modules shims of
`oracle/ref_shim.py`, builds seeded cluster states as plain kube-API style
dicts (the same dicts the product's host layer ingests), drives the reference's
own entry points and records what they did:

  * the reference's known-answer tests re-stated on the same values
    (test/test_cluster.py:56-73, test/test_scaler.py:37-77),
  * BASELINE config 1 (16 pending busybox pods x 4 D2_v2 nodes, dry-run loop_logic),
  * seeded random ticks (multi-dimension requests, several pools, running pods,
    DaemonSet / RC / bare pods, cordoned nodes, ignored pools, duplicate uids,
    the max_size raise path, the " pods" typo type, nvidia.com/gpu pods ...),
  * parse_SI / parse_resource vectors (utils.py:33-49).

The outputs are the committed fixtures the oracle (oracle/acsfit_oracle.c), the
host mirror and the CUDA path are all checked against.  `/root/reference` does
not exist on the GPU box; nothing there reads it.
"""
import copy
import datetime
import gzip
import json
import logging
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

OUT_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden")
NOW = datetime.datetime(2026, 9, 21, 12, 0, 0, tzinfo=datetime.timezone.utc)

GPU_KEY = "alpha.kubernetes.io/nvidia-gpu"


# ----------------------------------------------------------------------------
# kube-API style object builders (plain dicts)
# ----------------------------------------------------------------------------
def iso(dt):
    return dt.strftime("%Y-%m-%dT%H:%M:%SZ")


def make_node(pool, index, instance_type, created, unschedulable=None, cordon_label=None,
              cluster_id="16334397"):
    name = "k8s-%s-%s-%d" % (pool, cluster_id, index)
    labels = {
        "beta.kubernetes.io/instance-type": instance_type,
        "failure-domain.beta.kubernetes.io/region": "southcentralus",
        "kubernetes.io/hostname": name,
    }
    if cordon_label is not None:
        labels["openai/cordoned-by-autoscaler"] = cordon_label
    spec = {}
    if unschedulable is not None:
        spec["unschedulable"] = unschedulable
    return {"metadata": {"name": name, "labels": labels, "creationTimestamp": iso(created)},
            "spec": spec}


def make_pod(name, uid, phase, containers, node_name=None, created=None, start=None,
             annotations=None, labels=None, namespace="default"):
    md = {"name": name, "namespace": namespace, "uid": uid,
          "creationTimestamp": iso(created or NOW - datetime.timedelta(hours=5))}
    if annotations:
        md["annotations"] = annotations
    if labels:
        md["labels"] = labels
    spec = {"containers": [{"name": "c%d" % i, "resources": ({"requests": r} if r is not None else {})}
                           for i, r in enumerate(containers)]}
    if node_name:
        spec["nodeName"] = node_name
    status = {"phase": phase}
    if start is not None:
        status["startTime"] = iso(start)
    return {"metadata": md, "spec": spec, "status": status}


def created_by(kind):
    return json.dumps({"kind": "SerializedReference", "apiVersion": "v1",
                       "reference": {"kind": kind, "namespace": "system", "name": "x"}})


# ----------------------------------------------------------------------------
# reference driver
# ----------------------------------------------------------------------------
class _FakeDatetimeModule(types.ModuleType):
    """replaces the `datetime` module object inside autoscaler.scaler / autoscaler.kube so
    that datetime.datetime.now(tz) is deterministic (scaler.py:78, kube.py:68)."""

    def __init__(self, now):
        super().__init__("datetime")
        real = datetime

        class FixedDateTime(real.datetime):
            @classmethod
            def now(cls, tz=None):
                return now.astimezone(tz) if tz is not None else now.replace(tzinfo=None)

        self.datetime = FixedDateTime
        self.timedelta = real.timedelta
        self.timezone = real.timezone


class _ListHandler(logging.Handler):
    def __init__(self):
        super().__init__(level=logging.DEBUG)
        self.records = []

    def emit(self, record):
        self.records.append([record.levelname, record.name, record.getMessage()])


def res_hex(kube_resource):
    return {k: float(v).hex() for k, v in kube_resource.raw.items()}


def run_tick(ns, state):
    """drive the reference's Cluster.loop_logic (cluster.py:135-182) on `state`."""
    st = state["settings"]
    journal = []
    pk = ns.pykube

    def node_obj(obj):
        o = pk.Node(None, copy.deepcopy(obj))
        o.reload = lambda: journal.append(["node.reload", o.name])
        o.update = lambda: journal.append(["node.update", o.name, bool(o.obj["spec"].get("unschedulable")),
                                           o.obj["metadata"]["labels"].get("openai/cordoned-by-autoscaler")])
        o.delete = lambda: journal.append(["node.delete", o.name])
        return o

    def pod_obj(obj):
        o = pk.Pod(None, copy.deepcopy(obj))
        o.delete = lambda: journal.append(["pod.delete", o.name])
        return o

    pk.Node.objects.items = [node_obj(o) for o in state["nodes"]]
    pk.Pod.objects.items = [pod_obj(o) for o in state["pods"]]

    fake_dt = _FakeDatetimeModule(NOW)
    ns.scaler.datetime = fake_dt
    ns.kube.datetime = fake_dt

    # pool.max_size overrides (agent_pool.py:18 hard-codes 100; it is a plain instance attribute)
    max_sizes = st.get("pool_max_size", {})
    orig_pool_init = ns.agent_pool.AgentPool.__init__

    def pool_init(self, pool_name, instance_type, nodes):
        orig_pool_init(self, pool_name, instance_type, nodes)
        if pool_name in max_sizes:
            self.max_size = max_sizes[pool_name]

    ns.agent_pool.AgentPool.__init__ = pool_init

    scale_calls = []
    orig_scale_pools = ns.engine_scaler.EngineScaler.scale_pools

    def scale_pools(self, new_pool_sizes):
        scale_calls.append(dict(new_pool_sizes))
        return orig_scale_pools(self, new_pool_sizes)

    ns.engine_scaler.EngineScaler.scale_pools = scale_pools

    state_calls = []
    orig_state = ns.scaler.Scaler.get_node_state

    def get_node_state(self, node, node_pods, pods_to_schedule):
        s = orig_state(self, node, node_pods, pods_to_schedule)
        state_calls.append([node.name, s])
        return s

    ns.scaler.Scaler.get_node_state = get_node_state

    deleted = []
    ns.engine_scaler.delete_resources_for_node = lambda node, rg: deleted.append(node.name)

    class Deployments(object):
        def __init__(self):
            self.requested_pool_sizes = None

        def deploy(self, func, new_pool_sizes):
            journal.append(["deploy", dict(new_pool_sizes)])

    notifier = None
    if st.get("notifier") == "recording":
        class RecNotifier(object):
            def notify_scale(self, new, pods, cur):
                journal.append(["notify_scale", dict(new), [p.name for p in pods], dict(cur)])

            def notify_failed_to_scale(self, *a):
                journal.append(["notify_failed_to_scale"])

            def notify_drained_node(self, node, pods):
                journal.append(["notify_drained_node", node.name, [p.name for p in pods]])
        notifier = RecNotifier()

    handler = _ListHandler()
    root = logging.getLogger("autoscaler")
    root.setLevel(logging.DEBUG)
    root.addHandler(handler)
    root.propagate = False

    cl = ns.cluster.Cluster(
        kubeconfig=None, idle_threshold=st["idle_threshold"], spare_agents=st["spare_agents"],
        service_principal_app_id="x", service_principal_secret="x", service_principal_tenant_id="x",
        subscription_id="x", client_private_key="x", ca_private_key="x", instance_init_time=600,
        resource_group="rg", notifier=notifier, ignore_pools=st["ignore_pools"],
        scale_up=st.get("scale_up", True), maintainance=st.get("maintainance", True),
        over_provision=st["over_provision"], dry_run=st["dry_run"])
    cl.deployments = Deployments()
    cl.api = None
    cl.arm_template = {}
    cl.arm_parameters = copy.deepcopy(state["arm_parameters"])

    # capture the KubeNode objects to read used_capacity afterwards
    kube_nodes = []
    orig_create = cl.create_kube_node

    def create_kube_node(node):
        kn = orig_create(node)
        kube_nodes.append(kn)
        return kn

    cl.create_kube_node = create_kube_node

    result = {}
    try:
        result["return"] = cl.loop_logic()
        result["exception"] = None
    except Exception as e:  # the reference raises deliberately-looking errors (scaler.py:179-181)
        result["return"] = None
        result["exception"] = [type(e).__name__, str(e)]
    finally:
        root.removeHandler(handler)
        ns.engine_scaler.EngineScaler.scale_pools = orig_scale_pools
        ns.scaler.Scaler.get_node_state = orig_state
        ns.agent_pool.AgentPool.__init__ = orig_pool_init

    result["log"] = handler.records
    result["scale_pools_calls"] = scale_calls
    result["state_calls"] = state_calls
    result["journal"] = journal
    result["deleted_nodes"] = sorted(deleted)
    result["used"] = {kn.name: res_hex(kn.used_capacity) for kn in kube_nodes if not ns.utils.is_master(kn)}
    return result


# ----------------------------------------------------------------------------
# states
# ----------------------------------------------------------------------------
def arm_params(pools):
    p = {"masterVMSize": {"value": "Standard_D2_v2"}, "firstConsecutiveStaticIP": {"value": "10.240.255.5"}}
    for name, itype in pools:
        p[name + "Count"] = {"value": 1}
        p[name + "VMSize"] = {"value": itype}
    return p


def default_settings(**over):
    s = {"idle_threshold": 1800, "spare_agents": 1, "over_provision": 0, "dry_run": True,
         "ignore_pools": "", "pool_max_size": {}}
    s.update(over)
    return s


def state_config1():
    """BASELINE.json configs[0]: 16 pending busybox-like pods x 4 D2_v2 nodes x 1 populated pool."""
    created = datetime.datetime(2016, 8, 25, 5, 13, 16, tzinfo=datetime.timezone.utc)
    nodes = [make_node("agentpool1", i, "Standard_D2_v2", created) for i in range(4)]
    pods = [make_pod("busybox-%d" % i, "uid-%04d" % i, "Pending", [{"cpu": "1500m"}],
                     created=datetime.datetime(2016, 7, 14, 6, 46, 14, tzinfo=datetime.timezone.utc))
            for i in range(16)]
    return {"name": "config1_16x4_dry_run",
            "arm_parameters": arm_params([("agentpool1", "Standard_D2_v2"), ("agentpool2", "Standard_NC6")]),
            "nodes": nodes, "pods": pods,
            "settings": default_settings(idle_threshold=60, spare_agents=1, over_provision=0)}


POOL_TYPES = ["Standard_D2_v2", "Standard_D4_v3", "Standard_D8s_v3", "Standard_NC6", "Standard_NC12",
              "Standard_E16_v3", "Standard_M128s", "Standard_A1", "Standard_NC24", "Standard_F4"]

CPU_CHOICES = ["100m", "250m", "500m", "1", "1500m", "2", "4", "300m", "750m", "50m", "7"]
MEM_CHOICES = ["64Mi", "128Mi", "256Mi", "512Mi", "1Gi", "4Gi", "1500M", "123456789", "2G", "100Ki", "30Gi"]


def random_state(seed, n_pools=None, n_nodes=None, n_pending=None, n_running=None, **settings):
    rng = np.random.Generator(np.random.PCG64(seed))
    n_pools = n_pools or int(rng.integers(1, 5))
    types_ = [POOL_TYPES[int(i)] for i in rng.choice(len(POOL_TYPES), size=n_pools, replace=False)]
    pools = [("pool%c" % (97 + i), t) for i, t in enumerate(types_)]
    n_nodes = n_nodes if n_nodes is not None else int(rng.integers(1, 13))
    nodes = []
    per_pool_idx = {}
    for _ in range(n_nodes):
        pool, itype = pools[int(rng.integers(0, n_pools))]
        idx = per_pool_idx.get(pool, 0)
        per_pool_idx[pool] = idx + 1
        age = [5, 59, 61, 1799, 1801, 4000, 86399, 86400 + 10, 3 * 86400 + 10, 3 * 86400 + 5000][int(rng.integers(0, 10))]
        unsched = [None, None, None, False, True][int(rng.integers(0, 5))]
        cordon = [None, "true", "false"][int(rng.integers(0, 3))] if unsched else None
        nodes.append(make_node(pool, idx, itype, NOW - datetime.timedelta(seconds=age), unsched, cordon))

    def containers():
        out = []
        for _ in range(int(rng.integers(1, 4))):
            r = {}
            if rng.random() < 0.85:
                r["cpu"] = CPU_CHOICES[int(rng.integers(0, len(CPU_CHOICES)))]
            if rng.random() < 0.8:
                r["memory"] = MEM_CHOICES[int(rng.integers(0, len(MEM_CHOICES)))]
            if rng.random() < 0.12:
                r[GPU_KEY] = str(int(rng.integers(1, 3)))
            if rng.random() < 0.03:
                r["nvidia.com/gpu"] = "1"
            if rng.random() < 0.05:
                r["ephemeral-storage"] = "1Gi"
            out.append(r if (r or rng.random() < 0.5) else None)
        return out

    pods = []
    n_pending = n_pending if n_pending is not None else int(rng.integers(0, 30))
    n_running = n_running if n_running is not None else int(rng.integers(0, 40))
    kinds = ["pending"] * n_pending + ["running"] * n_running
    rng.shuffle(kinds)
    node_names = [n["metadata"]["name"] for n in nodes] + ["k8s-ghost-16334397-9"]
    for i, kind in enumerate(kinds):
        ann = {}
        r = rng.random()
        if r < 0.25:
            ann["kubernetes.io/created-by"] = created_by("DaemonSet")
        elif r < 0.7:
            ann["kubernetes.io/created-by"] = created_by("ReplicationController")
        if rng.random() < 0.05:
            ann["kubernetes.io/config.mirror"] = "abc"
        labels = {}
        if rng.random() < 0.1:
            labels["openai/do-not-drain"] = ["true", "1", "false", "True"][int(rng.integers(0, 4))]
        name = ("kube-proxy-%d" if rng.random() < 0.08 else "pod-%d") % i
        start = None
        if rng.random() < 0.85:
            start = NOW - datetime.timedelta(seconds=int([10, 3599, 3600, 3601, 86400 * 2][int(rng.integers(0, 5))]))
        if kind == "pending":
            pods.append(make_pod(name, "uid-%d" % i, "Pending", containers(), annotations=ann, labels=labels))
        else:
            phase = ["Running", "Running", "Running", "ContainerCreating", "Pending", "Succeeded", "Failed"][int(rng.integers(0, 7))]
            nn = node_names[int(rng.integers(0, len(node_names)))]
            pods.append(make_pod(name, "uid-%d" % i, phase, containers(), node_name=nn, start=start,
                                 annotations=ann, labels=labels))
    st = default_settings(
        idle_threshold=int([0, 60, 1800, 4000, 90000][int(rng.integers(0, 5))]),
        spare_agents=int(rng.integers(0, 3)),
        over_provision=int(rng.integers(0, 3)),
        dry_run=bool(rng.random() < 0.6),
        ignore_pools=["", "", pools[0][0], "nosuchpool," + pools[-1][0]][int(rng.integers(0, 4))],
        pool_max_size={p: int(rng.integers(1, 30)) for p, _ in pools if rng.random() < 0.3})
    st.update(settings)
    return {"name": "random_%d" % seed, "arm_parameters": arm_params(pools), "nodes": nodes, "pods": pods,
            "settings": st}


def crowded_state(seed, n_nodes=24, n_pending=160, n_running=120):
    """a state shaped like the benchmark: many pending pods, mostly-full nodes, max_size lifted."""
    s = random_state(seed, n_pools=3, n_nodes=n_nodes, n_pending=n_pending, n_running=n_running,
                     ignore_pools="", dry_run=True, over_provision=1)
    s["name"] = "crowded_%d" % seed
    pools = [k[:-6] for k in s["arm_parameters"] if k.endswith("VMSize") and k != "masterVMSize"]
    s["settings"]["pool_max_size"] = {p: 100000 for p in pools}
    return s


def duplicate_uid_state(ns):
    """two pods of the final pending list share a uid: the accounted dict collapses them but
    num_unaccounted = len(pods) still counts both (scaler.py:119-120) => the raise path."""
    s = random_state(777, n_pools=2, n_nodes=3, n_pending=8, n_running=4, ignore_pools="", dry_run=True)
    s["name"] = "duplicate_uids"
    s["settings"]["pool_max_size"] = {}
    probe = run_tick(ns, copy.deepcopy(s))
    pending_names = [m for lvl, lg, m in probe["log"] if lvl == "DEBUG" and lg == "autoscaler.cluster"]
    by_name = {p["metadata"]["name"]: p for p in s["pods"]}
    by_name[pending_names[3]]["metadata"]["uid"] = by_name[pending_names[1]]["metadata"]["uid"]
    return s


# ----------------------------------------------------------------------------
# function-level known-answer tests of the reference (same values as its own tests)
# ----------------------------------------------------------------------------
def kat_vectors(ns):
    pk = ns.pykube
    created = datetime.datetime(2016, 8, 25, 5, 13, 16, tzinfo=datetime.timezone.utc)
    out = {}

    # test/test_cluster.py:56-73
    node_o = make_node("agentpool1", 0, "Standard_D2_v2", created)
    pod_o = make_pod("busybox", "a85c73b6", "Running", [{"cpu": "1500m"}], node_name="10.0.0.228")
    cl = ns.cluster.Cluster(kubeconfig="~/.kube/config", idle_threshold=60, spare_agents=1,
                            instance_init_time=60, resource_group="my-rg", notifier=None,
                            service_principal_app_id="d", service_principal_secret="d",
                            service_principal_tenant_id="d", subscription_id="d",
                            client_private_key="d", ca_private_key="d", ignore_pools="", over_provision=0)

    def fresh_node():
        n = ns.kube.KubeNode(pk.Node(None, copy.deepcopy(node_o)))
        n.capacity = ns.capacity.get_capacity_for_instance_type(n.instance_type)
        return n

    pod = ns.kube.KubePod(pk.Pod(None, copy.deepcopy(pod_o)))
    n1 = fresh_node()
    a = cl.get_pending_pods([pod], [n1])
    n2 = fresh_node()
    b = cl.get_pending_pods([pod, ns.kube.KubePod(pk.Pod(None, copy.deepcopy(pod_o))),
                             ns.kube.KubePod(pk.Pod(None, copy.deepcopy(pod_o)))], [n2])
    out["test_get_pending_pods"] = {
        "node": node_o, "pod": pod_o,
        "pending_counts": [len(a), len(b)],
        "used_after": [res_hex(n1.used_capacity), res_hex(n2.used_capacity)]}

    # test/test_scaler.py:37-77  (node names there use the prefix 'k8-')
    def nodes(nb_pool, per_pool):
        res = []
        for p in range(nb_pool):
            for i in range(per_pool):
                o = make_node("agentpool%d" % (p + 1), i, "Standard_D2_v2", created)
                o["metadata"]["name"] = "k8-agentpool%d-16334397-%d" % (p + 1, i)
                res.append(o)
        return res

    params = arm_params([("agentpool1", "Standard_D2_v2"), ("agentpool2", "Standard_NC6")])

    def scaler_for(node_objs):
        kn = [ns.kube.KubeNode(pk.Node(None, copy.deepcopy(o))) for o in node_objs]
        return ns.engine_scaler.EngineScaler(
            resource_group="my-rg", nodes=kn, deployments=None, dry_run=False, over_provision=0,
            spare_count=1, arm_parameters=copy.deepcopy(params), arm_template={}, ignore_pools="",
            idle_threshold=0, notifier="")

    sc = scaler_for(nodes(2, 1))
    pools_a = [[p.name, p.instance_type, p.actual_capacity] for p in sc.agent_pools]
    kn3 = [ns.kube.KubeNode(pk.Node(None, copy.deepcopy(o))) for o in nodes(2, 3)]
    pools_b, _ = sc.get_agent_pools(kn3)
    out["test_get_agent_pools"] = {"arm_parameters": params, "nodes_2x1": nodes(2, 1), "nodes_2x3": nodes(2, 3),
                                   "pools_2x1": pools_a,
                                   "pools_2x3": [[p.name, p.instance_type, p.actual_capacity] for p in pools_b]}

    calls = []
    sc.scale_pools = lambda sizes: calls.append(dict(sizes))
    pod_b = copy.deepcopy(pod_o)
    p1 = ns.kube.KubePod(pk.Pod(None, copy.deepcopy(pod_b)))
    sc.fulfill_pending([p1])
    pod2 = copy.deepcopy(pod_o)
    pod2["spec"]["containers"][0]["resources"]["requests"]["cpu"] = "400m"
    pod2["metadata"]["uid"] = "fake"
    p2 = ns.kube.KubePod(pk.Pod(None, copy.deepcopy(pod2)))
    sc.fulfill_pending([p1, p2])
    pod3 = copy.deepcopy(pod2)
    pod3["spec"]["containers"][0]["resources"]["requests"]["cpu"] = "600m"
    p3 = ns.kube.KubePod(pk.Pod(None, copy.deepcopy(pod3)))
    sc.fulfill_pending([p1, p3])
    out["test_fulfill_pending"] = {"arm_parameters": params, "nodes": nodes(2, 1),
                                   "pod_1500m": pod_o, "pod_400m": pod2, "pod_600m": pod3,
                                   "scale_pools_calls": calls}
    return out


def parse_vectors(ns):
    samples = ["100m", "1500m", "1", "2", "0", "3952Mi", "64Mi", "1Gi", "1G", "12Ki", "5k", "7M", "3T",
               "1P", "2E", "9Ti", "4Pi", "1Ei", "250u", "3n", "17p", "0.5", "1.5Gi", "1e3", "abc", "",
               "12x", "007", "100mm", "123456789012345678901234567890", "10d", "10c", "1Z", "1Y", "1y",
               "1z", "1a", "1f", "-5", "+5", " 5", "5 "]
    out = []
    for s in samples:
        rec = {"s": s}
        try:
            rec["parse_SI"] = ns.utils.parse_SI(s).hex()
        except Exception as e:
            rec["parse_SI_error"] = type(e).__name__
        try:
            rec["parse_resource"] = float(ns.utils.parse_resource(s)).hex()
        except Exception as e:
            rec["parse_resource_error"] = type(e).__name__
        out.append(rec)
    bools = [None, "true", "True", "TRUE", "1", "0", "false", "yes", 1, True, False, ""]
    return {"quantities": out,
            "bool_labels": [[b, ns.utils.parse_bool_label(b)] for b in bools]}


def capacity_table(ns):
    """the instance-type table exactly as the reference loads it (capacity.py:12-18), in
    file order (= cost order, capacity.py:34-36), values as hex floats."""
    rows = []
    for itype in ns.capacity.data.keys():
        rows.append([itype, res_hex(ns.capacity.RESOURCE_SPEC[itype])])
    return rows


def main():
    ns = ref_shim.load_reference()
    os.makedirs(OUT_DIR, exist_ok=True)

    def dump(name, obj):
        path = os.path.join(OUT_DIR, name)
        text = json.dumps(obj, indent=None, separators=(",", ":"), sort_keys=False) + "\n"
        if name.endswith(".gz"):
            with open(path, "wb") as raw, gzip.GzipFile(filename="", mode="wb", fileobj=raw, mtime=0) as f:
                f.write(text.encode())
        else:
            with open(path, "w") as f:
                f.write(text)
        print("wrote %s (%d bytes)" % (path, os.path.getsize(path)))

    dump("capacity_table.json", {"source": "reference capacity.RESOURCE_SPEC (data/capacity.json order)",
                                 "rows": capacity_table(ns)})
    dump("parse_vectors.json", parse_vectors(ns))
    dump("kat_reference_tests.json", kat_vectors(ns))

    ticks = [state_config1()]
    ticks += [random_state(1000 + i) for i in range(36)]
    ticks += [crowded_state(2000 + i) for i in range(4)]
    ticks += [duplicate_uid_state(ns)]
    # the failure path with a notifier that has the attribute: NameError on selectors_hash (scaler.py:181)
    s = random_state(1007, notifier="recording")
    s["name"] = "random_1007_recording_notifier"
    ticks.append(s)
    s = crowded_state(2001)
    s["name"] = "crowded_2001_recording_notifier"
    s["settings"]["notifier"] = "recording"
    s["settings"]["dry_run"] = False
    ticks.append(s)
    # scale_up off / maintenance off
    s = random_state(1005, scale_up=False)
    s["name"] = "random_1005_no_scale"
    ticks.append(s)
    s = random_state(1006, maintainance=False)
    s["name"] = "random_1006_no_maintenance"
    ticks.append(s)
    # empty node list: loop_logic returns False (cluster.py:137-140)
    s = random_state(1007, n_nodes=0)
    s["name"] = "no_nodes"
    ticks.append(s)

    out = []
    for st in ticks:
        st = copy.deepcopy(st)
        st["now"] = iso(NOW)
        st["expected"] = run_tick(ns, st)
        out.append(st)
        exp = st["expected"]
        print("  %-36s ret=%s exc=%s scale_calls=%d states=%d" % (
            st["name"], exp["return"], exp["exception"] and exp["exception"][0],
            len(exp["scale_pools_calls"]), len(exp["state_calls"])))
    dump("ticks.json.gz", out)


if __name__ == "__main__":
    main()
