"""Host-layer behaviour at the kube / cloud adapter boundary (no GPU): the process-wide pykube settings the
reference applies at import (cluster.py:20-28), the secure-string placeholders (cluster.py:91-109) and the
default capacity table (config.py:5, capacity.py:12-18)."""
import sys
import types

import pytest


class _Objects(object):
    """pykube-like query descriptor whose DEFAULT namespace is 'default' (as pykube's ObjectManager)."""
    namespace = "default"

    def __init__(self, store):
        self.store = store

    def __call__(self, api):
        ns = self.namespace
        return [o for o in self.store if ns is None or o["metadata"].get("namespace") == ns]


def _fake_pykube(pods):
    mod = types.ModuleType("pykube")
    pod_objects = _Objects(pods)
    mod.Pod = type("Pod", (), {"objects": pod_objects})
    mod.Node = type("Node", (), {"objects": _Objects([])})
    conn = types.SimpleNamespace(match_hostname=None)
    mod.http = types.SimpleNamespace(requests=types.SimpleNamespace(
        packages=types.SimpleNamespace(urllib3=types.SimpleNamespace(connection=conn))))
    return mod


def test_list_pods_covers_all_namespaces(monkeypatch):
    from kubernetes_acs_engine_autoscaler_b200 import cluster as cl
    pods = [{"metadata": {"name": "a", "namespace": "default"}},
            {"metadata": {"name": "kube-dns", "namespace": "kube-system"}}]
    fake = _fake_pykube(pods)
    monkeypatch.setitem(sys.modules, "pykube", fake)
    monkeypatch.setattr(cl, "_pykube_ready", False)
    c = cl.Cluster(None, 1800, 1, "id", "secret", "tenant", "sub", "ckey", "cakey", 600, "rg", None, "")
    c.api = object()
    names = [p["metadata"]["name"] for p in c.list_pods()]
    assert names == ["a", "kube-dns"]           # a 'default'-only listing would drop the system pod
    assert fake.Pod.objects.namespace is None   # the setting the reference makes at import


def test_secure_string_placeholders_are_not_secrets(monkeypatch):
    from kubernetes_acs_engine_autoscaler_b200 import adapters
    from kubernetes_acs_engine_autoscaler_b200 import cluster as cl
    monkeypatch.setitem(adapters._overrides, "delete_master_vm_extension", lambda t: t)
    monkeypatch.delenv("ACSFIT_PLACEHOLDER_KEY", raising=False)
    c = cl.Cluster(None, 1800, 1, "app", "SECRET", "tenant", "sub", "CLIENT-KEY", "CA-KEY", 600, "rg", None, "")
    c.arm_parameters = {"etcdPeerPrivateKey0": {"value": "x"}, "etcdPeerPrivateKey3": {"value": "y"}}
    c.arm_template = {}
    c.fill_parameters_secure_strings()
    p = c.arm_parameters
    assert p["clientPrivateKey"]["value"] == "CLIENT-KEY" and p["caPrivateKey"]["value"] == "CA-KEY"
    assert p["servicePrincipalClientSecret"]["value"] == "SECRET"
    for key in ("kubeConfigPrivateKey", "apiServerPrivateKey", "etcdClientPrivateKey", "etcdServerPrivateKey",
                "etcdPeerPrivateKey0", "etcdPeerPrivateKey3"):
        assert p[key]["value"] == cl.PLACEHOLDER_KEY
        assert p[key]["value"] not in ("CA-KEY", "CLIENT-KEY", "SECRET", None)
    assert "etcdPeerPrivateKey1" not in p  # only the peers the deployment has (cluster.py:103-106)


def test_default_capacity_table_is_shipped():
    """Config.CAPACITY_DATA defaults to a file that exists, with the reference's key order and quirks."""
    import json
    import os
    from collections import OrderedDict
    from kubernetes_acs_engine_autoscaler_b200 import config
    path = config._default_capacity_data()
    assert os.path.exists(path)
    with open(path) as f:
        table = json.load(f, object_pairs_hook=OrderedDict)
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "golden", "capacity_table.json")) as f:
        rows = json.load(f, object_pairs_hook=OrderedDict)["rows"]
    assert list(table.keys()) == [r[0] for r in rows]
    for name, amounts in rows:
        assert list(table[name].items()) == [(k, float.fromhex(v)) for k, v in amounts.items()]
    assert " pods" in table["Standard_M128s"]  # the mistyped key of the last row changes results: kept


def test_missing_capacity_table_fails_at_load(tmp_path):
    from kubernetes_acs_engine_autoscaler_b200 import capacity
    with pytest.raises(FileNotFoundError):
        capacity.load(str(tmp_path / "nope.json"))
    capacity.load()  # restore the default table for the other tests
