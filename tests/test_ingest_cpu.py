"""`-m "not gpu"`: host-side ingestion helpers that replace a slow library call on the path must return
exactly what the reference's call returns (value, zone behaviour, exception class)."""
import random

import pytest
from dateutil.parser import parse as dateutil_parse

from kubernetes_acs_engine_autoscaler_b200 import utils


def _outcome(fn, text):
    try:
        v = fn(text)
        return ("ok", v, v.utcoffset(), type(v.tzinfo).__name__ if v.tzinfo else None)
    except Exception as e:  # noqa: BLE001 - the class is what is compared
        return ("raise", type(e).__name__)


CASES = [
    "2017-08-01T12:34:56Z", "2016-02-29T23:59:59Z", "1970-01-01T00:00:00Z", "9999-12-31T23:59:59Z", "0001-01-01T00:00:00Z",
    # out-of-range fields and other layouts must take dateutil's own route (same result or same error)
    "2017-02-30T00:00:00Z", "2017-13-01T00:00:00Z", "2017-08-01T24:00:00Z", "2017-08-01T12:60:00Z", "2017-08-01T12:34:60Z",
    "2017-08-01T12:34:56.5Z", "2017-08-01T12:34:56.123456Z", "2017-08-01T12:34:56+02:00", "2017-08-01 12:34:56",
    "2017-08-01T12:34:56z", "2017-08-01T12:34:56Z ", "20170801T123456Z", "Tue, 01 Aug 2017 12:34:56 GMT", "not a time", "",
]


@pytest.mark.parametrize("text", CASES)
def test_parse_time_matches_dateutil(text):
    assert _outcome(utils.parse_time, text) == _outcome(dateutil_parse, text)


def test_parse_time_matches_dateutil_random():
    rng = random.Random(20260921)
    for _ in range(2000):
        text = "%04d-%02d-%02dT%02d:%02d:%02dZ" % (rng.randint(1, 9999), rng.randint(1, 12), rng.randint(1, 31),
                                                   rng.randint(0, 23), rng.randint(0, 59), rng.randint(0, 59))
        assert _outcome(utils.parse_time, text) == _outcome(dateutil_parse, text)


def test_kubepod_times_are_the_references():
    from kubernetes_acs_engine_autoscaler_b200 import kube

    class Obj(object):
        def __init__(self, obj):
            self.obj = obj

    pod = kube.KubePod(Obj({"metadata": {"name": "p", "namespace": "d", "uid": "u", "creationTimestamp": "2017-08-01T12:34:56Z"},
                            "spec": {"containers": []}, "status": {"phase": "Pending", "startTime": "2017-08-01T12:35:00Z"}}))
    assert pod.creation_time == dateutil_parse("2017-08-01T12:34:56Z")
    assert pod.start_time == dateutil_parse("2017-08-01T12:35:00Z")
    assert type(pod.start_time.tzinfo) is type(dateutil_parse("2017-08-01T12:35:00Z").tzinfo)
