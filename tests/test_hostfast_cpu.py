"""The CPython helper module (csrc/hostfast.c) against the Python definitions it accelerates: KubePod construction
attribute for attribute (and exception for exception), identity grouping."""
import copy

import numpy as np
import pytest

from kubernetes_acs_engine_autoscaler_b200 import build as acs_build
from kubernetes_acs_engine_autoscaler_b200 import kube, snapshot
from kubernetes_acs_engine_autoscaler_b200 import synthetic as syn


@pytest.fixture(scope="module")
def hostfast():
    acs_build.build_hostfast()
    import importlib
    mod = importlib.import_module("kubernetes_acs_engine_autoscaler_b200._hostfast")
    return mod


class Raw(object):
    def __init__(self, obj):
        self.obj = obj


def both(hostfast, raws):
    """(python result or exception, C result or exception) for a list of raw pods"""
    def run(f):
        try:
            return f()
        except Exception as e:  # noqa: BLE001 - the exception itself is the result under test
            return e
    py = run(lambda: [kube.KubePod(r) for r in raws])
    c = run(lambda: hostfast.make_pods(kube.KubePod, raws, kube._TIME_MEMO, kube._remember_time, kube._RESOURCES_MEMO,
                                       kube._pod_resources))
    return py, c


def same(py, c):
    if isinstance(py, Exception) or isinstance(c, Exception):
        assert type(py) is type(c), (py, c)
        assert str(py) == str(c)
        return
    assert len(py) == len(c)
    for a, b in zip(py, c):
        assert type(b) is kube.KubePod
        assert list(a.__dict__) == list(b.__dict__)          # same attributes, same order
        for k in a.__dict__:
            va, vb = a.__dict__[k], b.__dict__[k]
            if k == "resources":
                assert va.raw == vb.raw
            else:
                assert va == vb, k
            assert type(va) is type(vb), k


def test_synthetic_cluster_pods_are_built_identically(hostfast):
    c = syn.make_cluster(3000, 300, 4, 2, seed=5)
    raws = [Raw(o) for o in syn.kube_objects(c)["pods"]]
    py, cc = both(hostfast, raws)
    same(py, cc)
    # the pods of one template share their KubeResource object in both routes (snapshot.Dims relies on it)
    assert len({id(p.resources) for p in cc}) == len({id(p.resources) for p in py})
    assert cc[0].original is raws[0]


def base_pod():
    return {"metadata": {"name": "p", "namespace": "ns", "uid": "u1", "creationTimestamp": "2017-09-01T10:00:00Z",
                         "labels": {"owner": "me"}, "annotations": {"a": "b"}},
            "spec": {"nodeName": "n1", "nodeSelector": {"k": "v"},
                     "containers": [{"resources": {"requests": {"cpu": "250m", "memory": "1Gi"}}}]},
            "status": {"phase": "Running", "startTime": "2017-09-01T10:00:05Z"}}


def mutations():
    def drop(path):
        def f(o):
            d = o
            for k in path[:-1]:
                d = d[k]
            del d[path[-1]]
        return f

    def put(path, value):
        def f(o):
            d = o
            for k in path[:-1]:
                d = d[k]
            d[path[-1]] = value
        return f
    yield "plain", lambda o: None
    for path in (("metadata",), ("spec",), ("status",), ("metadata", "name"), ("metadata", "namespace"), ("metadata", "uid"),
                 ("metadata", "creationTimestamp"), ("metadata", "labels"), ("metadata", "annotations"),
                 ("spec", "nodeName"), ("spec", "nodeSelector"), ("spec", "containers"), ("status", "phase"),
                 ("status", "startTime")):
        yield "drop " + "/".join(path), drop(path)
    yield "labels None", put(("metadata", "labels"), None)
    yield "bad creation text", put(("metadata", "creationTimestamp"), "not a time")
    yield "bad start text", put(("status", "startTime"), "yesterday-ish")
    yield "odd time format", put(("metadata", "creationTimestamp"), "2017-09-01 10:00:00+02:00")
    yield "two containers", put(("spec", "containers"), [{"resources": {"requests": {"cpu": "1"}}},
                                                          {"resources": {"requests": {"cpu": "2", "memory": "1Mi"}}}])
    yield "no containers", put(("spec", "containers"), [])
    yield "container without resources", put(("spec", "containers"), [{}])
    yield "resources None", put(("spec", "containers"), [{"resources": None}])
    yield "empty resources", put(("spec", "containers"), [{"resources": {}}])
    yield "requests None", put(("spec", "containers"), [{"resources": {"requests": None}}])
    yield "empty requests", put(("spec", "containers"), [{"resources": {"requests": {}}}])
    yield "bad quantity", put(("spec", "containers"), [{"resources": {"requests": {"cpu": "lots"}}}])
    yield "container not a dict", put(("spec", "containers"), ["x"])
    yield "containers a tuple", put(("spec", "containers"), ({"resources": {"requests": {"cpu": "1"}}},))
    yield "unhashable quantity", put(("spec", "containers"), [{"resources": {"requests": {"cpu": ["1"]}}}])
    yield "metadata not a dict", put(("metadata",), "zzz")


@pytest.mark.parametrize("name,mutate", list(mutations()), ids=[n for n, _ in mutations()])
def test_odd_and_malformed_pods_behave_like_the_python_constructor(hostfast, name, mutate):
    good = base_pod()
    bad = copy.deepcopy(good)
    mutate(bad)
    # a good pod first (fills the memos), then the mutated one in the middle of a batch
    raws = [Raw(copy.deepcopy(good)), Raw(bad), Raw(copy.deepcopy(good))]
    py, cc = both(hostfast, raws)
    same(py, cc)


def test_obj_that_is_not_a_plain_dict_goes_through_the_constructor(hostfast):
    class D(dict):
        pass
    raws = [Raw(D(base_pod())), Raw(base_pod())]
    py, cc = both(hostfast, raws)
    same(py, cc)

    class NoObj(object):
        pass
    py, cc = both(hostfast, [NoObj()])
    same(py, cc)


def test_group_ids_matches_first_occurrence_grouping(hostfast):
    rng = np.random.default_rng(3)
    objs = [object() for _ in range(5000)]
    seq = [objs[i] for i in rng.integers(0, 5000, size=40000)]
    inv = np.empty(len(seq), dtype=np.int64)
    uniq = hostfast.group_ids(seq, inv)
    slot, ref_uniq, ref_inv = {}, [], []
    for o in seq:
        if id(o) not in slot:
            slot[id(o)] = len(ref_uniq)
            ref_uniq.append(o)
        ref_inv.append(slot[id(o)])
    assert all(a is b for a, b in zip(uniq, ref_uniq)) and len(uniq) == len(ref_uniq)
    np.testing.assert_array_equal(inv, np.asarray(ref_inv))
    assert hostfast.group_ids([], np.empty(0, dtype=np.int64)) == []
    with pytest.raises(ValueError):
        hostfast.group_ids(seq, np.empty(3, dtype=np.int64))


def test_make_pods_wrapper_falls_back_when_the_constructor_is_replaced(hostfast, monkeypatch):
    raws = [Raw(base_pod())]
    seen = []
    orig = kube.KubePod.__init__

    def init(self, pod):
        seen.append(pod)
        orig(self, pod)
    monkeypatch.setattr(kube.KubePod, "__init__", init)
    pods = kube.make_pods(raws)
    assert seen == raws and pods[0].name == "p"


def test_snapshot_grouping_without_the_helper(monkeypatch):
    a, b = object(), object()
    with_helper = snapshot._group([a, b, a, a, b])
    monkeypatch.setattr(snapshot, "_hostfast", None)
    plain = snapshot._group([a, b, a, a, b])
    np.testing.assert_array_equal(with_helper[0], plain[0])
    assert [x is y for x, y in zip(with_helper[1], plain[1])] == [True, True]
