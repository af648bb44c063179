"""Host-layer helpers of the flattening (snapshot.py) against their plain definitions."""
import datetime

import numpy as np

from kubernetes_acs_engine_autoscaler_b200 import snapshot, utils
from kubernetes_acs_engine_autoscaler_b200.kube import KubeResource


def test_touched_keys_is_the_union_of_the_counted_vectors_key_sets():
    rng = np.random.default_rng(11)
    vectors = [KubeResource(**{k: 1.0 for k in keys}) for keys in
               (("cpu",), ("cpu", "memory"), ("pods",), ("cpu", "alpha.kubernetes.io/nvidia-gpu"), ())]
    pods = [vectors[i] for i in rng.integers(0, len(vectors), size=4000)]
    node_of = rng.integers(0, 300, size=4000)
    inv, uniq = snapshot._group(pods)
    got = snapshot._touched_keys(node_of, inv, uniq)
    want = {}
    for n, v in zip(node_of.tolist(), pods):
        want.setdefault(n, set()).update(v.raw)
    assert got == want
    assert snapshot._touched_keys(np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64), []) == {}


def test_rows_with_groups_returns_the_same_rows():
    a, b = KubeResource(cpu=1.5, memory=2.0), KubeResource(pods=1.0)
    dims = snapshot.Dims([a, b])
    plain = dims.rows([a, b, a, a])
    rows, inv, uniq = dims.rows([a, b, a, a], groups=True)
    np.testing.assert_array_equal(rows, plain)
    assert inv.tolist() == [0, 1, 0, 0] and uniq[0] is a and uniq[1] is b


def test_node_age_with_a_shared_clock_reads_the_clock_once_per_zone(monkeypatch):
    tz = datetime.timezone.utc
    calls = []

    def fake_now(zone=None):
        calls.append(zone)
        return datetime.datetime(2017, 9, 2, 12, 0, 0, tzinfo=zone)
    monkeypatch.setattr(utils, "now", fake_now)

    class N(object):
        def __init__(self, t):
            self.creation_time = t
    nodes = [N(datetime.datetime(2017, 9, 1, 11, 0, 0, tzinfo=tz)), N(datetime.datetime(2017, 9, 2, 11, 59, 30, tzinfo=tz))]
    alone = [snapshot.node_age_seconds(n) for n in nodes]
    clock = {}
    shared = [snapshot.node_age_seconds(n, clock) for n in nodes]
    assert alone == shared == [3600, 30]          # .seconds wraps at one day, like the reference (scaler.py:78)
    assert len(calls) == 3                        # two single readings + ONE for the batch
