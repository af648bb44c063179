"""TEST INFRASTRUCTURE -- drives the UNMODIFIED reference (oracle/ref_shim.py) on the synthetic snapshots of
kubernetes_acs_engine_autoscaler_b200/synthetic.py.

Three users, none of them product code:
* bench.py's cpu_baseline leg: the reference's own CPython loop timed on scaled-down shapes of the benchmark
  workload, in the same run as the GPU numbers;
* oracle/make_golden.py: mid-size recorded ticks (placements, node state bits, pool sizes) for tests/golden/;
* tests/test_reference_live_cpu.py: differential checks against the C oracle wherever the reference is present.

A dense synthetic cluster is turned into kube-API style objects whose quantity STRINGS parse
(utils.parse_SI, reference utils.py:36-42) to exactly the float64 values of the dense rows; that the reference's
constructors give the rows back bit for bit is itself one of the tests.
"""
import datetime
import json
import logging
import os
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

import ref_shim  # noqa: E402
from kubernetes_acs_engine_autoscaler_b200 import synthetic as syn  # noqa: E402



# the kube-API style view of a synthetic snapshot lives with the generator (product side, also used by bench.py)
quantity = syn.quantity
requests_of = syn.requests_of
pool_name = syn.pool_name
node_name = syn.node_name
kube_state = syn.kube_objects
NOW = syn.NOW


def capacity_file(c, directory):
    """a CAPACITY_DATA file (config.py:5) holding the synthetic instance types in cost order, with the extra
    resources of the D = 8 shapes; D = 4 values equal the reference's own data/capacity.json."""
    table = {}
    for t in range(len(syn.INSTANCE_TYPES)):
        row = syn.capacity_rows(len(syn.INSTANCE_TYPES), c["D"])[t]
        table[syn.INSTANCE_TYPES[t][0]] = {name: float(v) for name, v in zip(c["dim_names"], row)}
    path = os.path.join(directory, "capacity_%d.json" % c["D"])
    with open(path, "w") as f:
        json.dump(table, f)
    return path


class ReferenceTick(object):
    """the reference's classes instantiated on one synthetic cluster, ready to run the path step by step."""

    def __init__(self, c, tmpdir=None, max_size=None):
        self._tmp = tmpdir or tempfile.mkdtemp(prefix="acsfit_ref_")
        self.ns = ref_shim.load_reference(capacity_data=capacity_file(c, self._tmp))
        ns, st = self.ns, kube_state(c)
        self.c = c
        pk = ns.pykube
        self.cl = ns.cluster.Cluster(
            kubeconfig=None, idle_threshold=1800, spare_agents=1, service_principal_app_id="x",
            service_principal_secret="x", service_principal_tenant_id="x", subscription_id="x",
            client_private_key="x", ca_private_key="x", instance_init_time=600, resource_group="rg",
            notifier=None, ignore_pools="", over_provision=c["over_provision"], dry_run=True)
        t0 = time.perf_counter()
        self.nodes = [self.cl.create_kube_node(pk.Node(None, o)) for o in st["nodes"]]
        self.pods = [ns.kube.KubePod(pk.Pod(None, o)) for o in st["pods"]]
        self.ingest_s = time.perf_counter() - t0
        self.scaler = ns.engine_scaler.EngineScaler(
            resource_group="rg", nodes=self.nodes, deployments=None, arm_template={},
            arm_parameters=st["arm_parameters"], dry_run=True, ignore_pools="",
            over_provision=c["over_provision"], spare_count=1, idle_threshold=1800, notifier=None)
        for pool in self.scaler.agent_pools:  # AgentPool.max_size is a plain attribute (agent_pool.py:18: 100)
            pool.max_size = int(max_size if max_size is not None else c["pool_max"][0])
        self.scale_calls = []
        self.scaler.scale_pools = lambda sizes: self.scale_calls.append(dict(sizes))
        self.running = [p for p in self.pods if p.status == "Running"]

    def occupancy(self):
        """cluster.py:159-168"""
        by_name = {n.name: n for n in self.nodes}
        for p in self.running:
            by_name[p.node_name].count_pod(p)  # (the reference scans all nodes per pod; same sums, same order)

    def node_states(self, pods_to_schedule):
        """Scaler.get_node_state for every node (scaler.py:61-114) with `now` pinned to NOW; state strings."""
        import make_golden
        fake = make_golden._FakeDatetimeModule(NOW)
        self.ns.scaler.datetime = fake
        self.ns.kube.datetime = fake
        by_node = {}
        for p in self.running:
            by_node.setdefault(p.node_name, []).append(p)
        return [self.scaler.get_node_state(n, by_node.get(n.name, []), pods_to_schedule) for n in self.nodes]

    def run(self, record=False):
        """get_pods_to_schedule + get_pending_pods + fulfill_pending, timed; returns a result dict.
        record=True also notes which node took which pod (KubeNode.count_pod is wrapped) and the node states."""
        cl, ns = self.cl, self.ns
        taken = {}
        if record:
            index_of = {n.name: i for i, n in enumerate(self.nodes)}
            orig_count = ns.kube.KubeNode.count_pod

            def count_pod(node, pod):
                if pod.status == "Pending":
                    taken[pod.uid] = index_of[node.name]
                return orig_count(node, pod)
            ns.kube.KubeNode.count_pod = count_pod
        logging.disable(logging.CRITICAL)
        try:
            self.occupancy()
            t0 = time.perf_counter()
            to_schedule = cl.get_pods_to_schedule(self.pods, self.scaler.agent_pools)
            t1 = time.perf_counter()
            pending = cl.get_pending_pods(to_schedule, self.nodes)
            t2 = time.perf_counter()
            exc = None
            if pending:
                try:
                    self.scaler.fulfill_pending(pending)
                except Exception as e:  # the reference's raise path (scaler.py:179-181)
                    exc = [type(e).__name__, str(e)]
            t3 = time.perf_counter()
        finally:
            logging.disable(logging.NOTSET)
            if record:
                ns.kube.KubeNode.count_pod = orig_count
        pend_uid = {p.uid for p in pending}
        sched_uid = [p.uid for p in to_schedule]
        names = self.c["dim_names"]
        used = np.array([[float(n.used_capacity.raw.get(k, 0.0)) for k in names] for n in self.nodes], dtype=np.float64)
        return {"seconds": {"get_pods_to_schedule": t1 - t0, "get_pending_pods": t2 - t1, "fulfill_pending": t3 - t2},
                "to_schedule": sched_uid, "pending": [u for u in sched_uid if u in pend_uid],
                "scale_calls": self.scale_calls, "exception": exc, "used": used, "ingest_seconds": self.ingest_s,
                "placed": [taken.get(u, -1) for u in sched_uid] if record else None,
                "states": self.node_states(to_schedule) if record else None}


def dense_answer(oracle, c):
    """the same tick on the dense layout by the C oracle: mask, placements, used, new sizes, decisions."""
    used = syn.initial_used(c)
    mask, ev0 = oracle.feasible_mask(c["req"], c["unit_all"])
    idx = np.nonzero(mask)[0]
    placed, ev1 = oracle.first_fit_nodes(c["req"][idx], c["cap_type"], c["node_type"], used)
    pend = idx[placed < 0]
    f = {"new_size": c["pool_actual"].astype(np.int64), "evals": 0, "num_unaccounted": 0}
    if len(pend):
        f = oracle.fulfill_pending(c["req"][pend], len(pend), c["unit_ordered"], c["pool_actual"], c["pool_max"],
                                   c["pool_ignored"], c["over_provision"])
    return {"feasible": idx, "placed": placed, "pending": pend, "used": used, "new_size": np.asarray(f["new_size"]),
            "num_unaccounted": int(f["num_unaccounted"]), "decisions": int(ev0 + ev1 + f["evals"]),
            "decisions_nodes": int(ev1), "decisions_bins": int(f["evals"])}


def compare(c, ref, dense):
    """reference result (ReferenceTick.run) against the dense answer; returns a list of mismatches."""
    bad = []
    if ref["to_schedule"] != ["pend-%d" % i for i in dense["feasible"]]:
        bad.append("pods to schedule")
    if ref["pending"] != ["pend-%d" % i for i in dense["pending"]]:
        bad.append("pending pods")
    if not np.array_equal(ref["used"].view(np.uint64), dense["used"].view(np.uint64)):
        bad.append("used bits")
    if dense["num_unaccounted"] == 0 and len(dense["pending"]):
        want = {pool_name(t): int(dense["new_size"][t]) for t in range(c["T"])}
        if ref["scale_calls"] != [want]:
            bad.append("scale_pools %r != %r" % (ref["scale_calls"], want))
    elif len(dense["pending"]) and ref["exception"] is None:
        bad.append("expected the raise path")
    return bad


def time_reference(shapes=((1000, 100), (2000, 200), (4000, 400)), D=4, T=1, seed=20260921 + 2, verify=True):
    """the reference's CPython loop on geometrically scaled-down shapes of the benchmark
    generator; decisions = the reference's own can_fit / .possible evaluation count (from the C oracle on the
    same data, against which the reference's results are also checked).  Single-threaded (GIL; order-dependent)."""
    import oracle
    oracle.build()
    rows, total_dec, total_s = [], 0, 0.0
    for (P, N) in shapes:
        c = syn.make_cluster(P, N, D, T, seed=seed)
        tick = ReferenceTick(c)
        res = tick.run()
        dense = dense_answer(oracle, c)
        mism = compare(c, res, dense) if verify else []
        s = sum(res["seconds"].values())
        rows.append({"pods": P, "nodes": N, "seconds": s, "decisions": dense["decisions"],
                     "decisions_per_s": dense["decisions"] / s, "pods_per_s": P / s,
                     "matches_oracle": not mism, "mismatches": mism, "parts": res["seconds"],
                     "ingest_seconds": res["ingest_seconds"]})
        total_dec += dense["decisions"]
        total_s += s
    return {"value": total_dec / total_s, "unit": "decisions/s", "cores": 1, "kind": "reference",
            "python": sys.version.split()[0], "shapes": rows,
            "sample": "the unmodified reference (oracle/_ref, CPython, 1 thread) on %s pods x nodes of the c2 generator: "
                      "get_pods_to_schedule + get_pending_pods + fulfill_pending, logging off"
                      % ", ".join("%dx%d" % s for s in shapes)}


if __name__ == "__main__":
    print(json.dumps(time_reference(), indent=1))
