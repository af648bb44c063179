"""`-m "not gpu"`: the reference's own recorded behaviour (tests/golden, generated from the UNMODIFIED
reference by oracle/make_golden.py) against (a) the plain-C oracle and (b) the product's host layer
(flattening, key bookkeeping, logs, errors, adapter calls) -- with the oracle standing in for the GPU."""
import copy
import json
import os

import pytest

import golden_util as gu


@pytest.fixture()
def host(tmp_path, monkeypatch, oracle_mod):
    from kubernetes_acs_engine_autoscaler_b200 import capacity, snapshot
    from oracle_engine import OracleEngine
    capacity.load(gu.write_capacity_file(tmp_path), cpu_reserve=0.0)
    monkeypatch.setattr(snapshot, "_engine", OracleEngine())
    return monkeypatch


TICKS = gu.load_json("ticks.json.gz")


@pytest.mark.parametrize("state", TICKS, ids=[s["name"] for s in TICKS])
def test_tick_matches_reference(host, state):
    got = gu.run_tick(state, host)
    gu.assert_tick_matches(got, state["expected"])


def test_config1_golden_shape():
    """BASELINE.json configs[0]: 16 pending x 4 nodes, dry run."""
    st = [s for s in TICKS if s["name"] == "config1_16x4_dry_run"][0]
    log = [m for _, _, m in st["expected"]["log"]]
    assert "Pods to schedule: 16" in log and "Pending pods: 12" in log
    assert "[Dry run] Would have scaled pool 'agentpool1' to 16 agent(s) (currently at 4)" in log
    assert "Pool 'agentpool2' already at desired capacity (0)" in log
    assert st["expected"]["scale_pools_calls"] == [{"agentpool1": 16, "agentpool2": 0}]
    assert [s for _, s in st["expected"]["state_calls"]] == ["pod-pending"] * 4


def test_reference_known_answer_tests(host):
    """the reference's own unit tests, on its own values (test/test_cluster.py:56-73,
    test/test_scaler.py:37-77), through the product's classes."""
    from kubernetes_acs_engine_autoscaler_b200 import capacity
    from kubernetes_acs_engine_autoscaler_b200.cluster import Cluster
    from kubernetes_acs_engine_autoscaler_b200.engine_scaler import EngineScaler
    from kubernetes_acs_engine_autoscaler_b200.kube import KubeNode, KubePod
    kat = gu.load_json("kat_reference_tests.json")
    j = []

    def node(obj):
        n = KubeNode(gu.FakeKubeObject(obj, j, "node"))
        n.capacity = capacity.get_capacity_for_instance_type(n.instance_type)
        return n

    def pod(obj):
        return KubePod(gu.FakeKubeObject(obj, j, "pod"))

    k = kat["test_get_pending_pods"]
    cl = Cluster(kubeconfig="~/.kube/config", idle_threshold=60, spare_agents=1, instance_init_time=60,
                 resource_group="my-rg", notifier=None, service_principal_app_id="d", service_principal_secret="d",
                 service_principal_tenant_id="d", subscription_id="d", client_private_key="d", ca_private_key="d",
                 ignore_pools="", over_provision=0)
    n1 = node(k["node"])
    a = cl.get_pending_pods([pod(k["pod"])], [n1])
    n2 = node(k["node"])
    b = cl.get_pending_pods([pod(k["pod"]), pod(k["pod"]), pod(k["pod"])], [n2])
    assert [len(a), len(b)] == k["pending_counts"] == [0, 2]
    for n, exp in zip((n1, n2), k["used_after"]):
        assert {key: float(v).hex() for key, v in n.used_capacity.raw.items()} == exp

    k = kat["test_get_agent_pools"]

    def scaler_for(node_objs, params):
        return EngineScaler(resource_group="my-rg", nodes=[KubeNode(gu.FakeKubeObject(o, j, "node")) for o in node_objs],
                            deployments=None, dry_run=False, over_provision=0, spare_count=1,
                            arm_parameters=copy.deepcopy(params), arm_template={}, ignore_pools="",
                            idle_threshold=0, notifier="")
    sc = scaler_for(k["nodes_2x1"], k["arm_parameters"])
    assert [[p.name, p.instance_type, p.actual_capacity] for p in sc.agent_pools] == k["pools_2x1"]
    pools_b, _ = sc.get_agent_pools([KubeNode(gu.FakeKubeObject(o, j, "node")) for o in k["nodes_2x3"]])
    assert [[p.name, p.instance_type, p.actual_capacity] for p in pools_b] == k["pools_2x3"]

    k = kat["test_fulfill_pending"]
    sc = scaler_for(k["nodes"], k["arm_parameters"])
    calls = []
    sc.scale_pools = lambda sizes: calls.append(dict(sizes))
    p1 = pod(k["pod_1500m"])
    sc.fulfill_pending([p1])
    sc.fulfill_pending([p1, pod(k["pod_400m"])])
    sc.fulfill_pending([p1, pod(k["pod_600m"])])
    assert calls == k["scale_pools_calls"] == [{"agentpool1": 2, "agentpool2": 1}, {"agentpool1": 2, "agentpool2": 1},
                                               {"agentpool1": 3, "agentpool2": 1}]


def test_parse_vectors():
    from kubernetes_acs_engine_autoscaler_b200 import utils
    vec = gu.load_json("parse_vectors.json")
    for rec in vec["quantities"]:
        s = rec["s"]
        for fn, key in ((utils.parse_SI, "parse_SI"), (utils.parse_resource, "parse_resource")):
            if key in rec:
                assert float(fn(s)).hex() == rec[key], (s, key)
            else:
                with pytest.raises(Exception) as ei:
                    fn(s)
                assert type(ei.value).__name__ == rec[key + "_error"], (s, key)
    for value, exp in vec["bool_labels"]:
        assert utils.parse_bool_label(value) == exp


def test_loop_swallows_errors_unless_debug(host):
    """Cluster.loop returns False on any exception, lets it through with debug (cluster.py:111-128)."""
    st = [s for s in TICKS if s["expected"]["exception"]][0]
    from kubernetes_acs_engine_autoscaler_b200 import cluster
    calls = {"n": 0}

    def boom(self):
        calls["n"] += 1
        raise AttributeError("x")
    host.setattr(cluster.Cluster, "loop_logic", boom)
    cl = cluster.Cluster(None, 1, 1, "a", "b", "c", "d", "e", "f", 1, "rg", None, "")
    assert cl.loop(False) is False and cl.scale_loop(False) is False
    with pytest.raises(AttributeError):
        cl.loop(True)
    assert calls["n"] == 3
