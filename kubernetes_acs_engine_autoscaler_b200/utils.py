"""Quantity parsing and node-name helpers -- host-side ingestion of the hot path.

These define the float64 BIT PATTERNS every later stage consumes, so they follow the reference
exactly (autoscaler/utils.py:6-74): a quantity is `float(digits) * multiplier`, never
`digits / 1000`.  tests/golden/parse_vectors.json pins them against the reference.
"""
import contextlib
import datetime
import gc
import functools
import re

_DECIMAL = [('y', 1e-24), ('z', 1e-21), ('a', 1e-18), ('f', 1e-15), ('p', 1e-12), ('n', 1e-9), ('u', 1e-6),
            ('m', 1e-3), ('c', 1e-2), ('d', 1e-1), ('k', 1e3), ('M', 1e6), ('G', 1e9), ('T', 1e12),
            ('P', 1e15), ('E', 1e18), ('Z', 1e21), ('Y', 1e24)]
_BINARY = [('Ki', 2 ** 10), ('Mi', 2 ** 20), ('Gi', 2 ** 30), ('Ti', 2 ** 40), ('Pi', 2 ** 50), ('Ei', 2 ** 60)]
# insertion order matters: the alternation below tries suffixes in this order (reference utils.py:33)
SI_suffix = dict(_DECIMAL + _BINARY)
SI_regex = re.compile(r"(\d+)(%s)?$" % "|".join(SI_suffix.keys()))


@functools.lru_cache(maxsize=1 << 16)
def _parse_SI_text(s):
    match = SI_regex.match(s)
    if match is None:
        raise ValueError("Unknown SI quantity: %s" % s)
    digits, suffix = match.groups()
    return float(digits) * (SI_suffix[suffix] if suffix else 1.)


def parse_SI(s):
    """'1500m' -> 1500.0 * 1e-3 ; '3952Mi' -> 3952.0 * 2**20 ; raises ValueError otherwise.
    A cluster holds few distinct quantity strings, so exact `str` inputs are memoised (a pure function of the
    text; errors are not cached and raise again); anything else takes the reference's route as is."""
    if type(s) is str:
        return _parse_SI_text(s)
    match = SI_regex.match(s)
    if match is None:
        raise ValueError("Unknown SI quantity: %s" % s)
    digits, suffix = match.groups()
    return float(digits) * (SI_suffix[suffix] if suffix else 1.)


def parse_resource(resource):
    try:
        return float(resource)
    except ValueError:
        return parse_SI(resource)


def parse_bool_label(value):
    return str(value).lower() in ('1', 'true')


def _name_parts(node):
    parts = node.name.split('-')
    if len(parts) != 4:
        raise ValueError('Kubernetes node name was malformed and cannot be processed.')
    return parts


def is_master(node):
    return _name_parts(node)[1] == 'master'


def is_agent(node):
    return not is_master(node)


def get_instance_index(node):
    return int(_name_parts(node)[3])


def get_pool_name(node):
    return _name_parts(node)[1]


_Z_ZONE = None
_RFC3339_Z = re.compile(r"(\d{4})-(\d\d)-(\d\d)T(\d\d):(\d\d):(\d\d)Z\Z")


def parse_time(text):
    """the timestamp the reference gets from dateutil.parser.parse (kube.py:37-38, :104): the API server's
    'YYYY-MM-DDTHH:MM:SSZ' form is decoded directly (dateutil spends ~40 us per call on it, which dominates
    ingestion at 10^5 pods); anything else, or any out-of-range field, goes through dateutil itself so the
    result and the errors are the reference's.  The zone object is the one dateutil itself attaches to a 'Z'
    timestamp on this host (tzlocal() where the local zone is UTC, tzutc() elsewhere), asked once.
    Exact `str` inputs are memoised: a pure function of the text, and pods created in one burst share theirs."""
    if type(text) is str:
        return _parse_time_text(text)
    from dateutil.parser import parse
    return parse(text)


@functools.lru_cache(maxsize=1 << 16)
def _parse_time_text(text):
    global _Z_ZONE
    if _Z_ZONE is None:
        from dateutil.parser import parse
        _Z_ZONE = parse("2000-01-01T00:00:00Z").tzinfo
    m = _RFC3339_Z.match(text)
    if m is not None:
        try:
            return datetime.datetime(int(text[0:4]), int(text[5:7]), int(text[8:10]), int(text[11:13]), int(text[14:16]),
                                     int(text[17:19]), tzinfo=_Z_ZONE)
        except ValueError:
            pass
    from dateutil.parser import parse
    return parse(text)


@contextlib.contextmanager
def gc_paused():
    """suspend the cyclic garbage collector while a tick builds its objects in bulk (nothing in a tick creates
    reference cycles that must be reclaimed mid-tick); restored on exit, whatever happens."""
    was = gc.isenabled()
    gc.disable()
    try:
        yield
    finally:
        if was:
            gc.enable()


# injectable clock: the reference calls datetime.datetime.now(tz) inline (scaler.py:78, kube.py:68);
# tests replace this hook to make node ages and drain grace periods deterministic.
def now(tz=None):
    return datetime.datetime.now(tz)
