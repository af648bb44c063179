"""Pod / node / resource-vector types of the tick (mirror of reference autoscaler/kube.py).

Same names, fields and error behaviour as the reference so that callers and tests read the
same.  The bulk arithmetic of a tick never goes through these Python objects: Cluster / Scaler
flatten them into dense float64 buffers (snapshot.py) and run the CUDA kernels.  The mutators
(cordon / uncordon / drain / delete) are the downward adapter boundary: they call straight
through to the wrapped pykube objects, exactly as the reference does (kube.py:124-167).
"""
import datetime
import json
import logging
import os

from . import utils

logger = logging.getLogger('autoscaler.kube')

try:  # CPython twin of the two hottest host loops (csrc/hostfast.c); the Python code below is the definition
    if os.environ.get('ACSFIT_NO_HOSTFAST'):
        raise ImportError('disabled')
    from . import _hostfast
except ImportError:
    _hostfast = None

_CORDON_LABEL = 'openai/cordoned-by-autoscaler'


class KubePodStatus(object):
    RUNNING = 'Running'
    PENDING = 'Pending'
    CONTAINER_CREATING = 'ContainerCreating'
    SUCCEEDED = 'Succeeded'
    FAILED = 'Failed'


class KubeResource(object):
    """sparse resource vector: dict key -> float64, absent key == 0 (kube.py:197-249)."""

    def __init__(self, **kwargs):
        self.raw = {key: utils.parse_resource(value) for key, value in kwargs.items()}

    def _zip(self, other, op):
        merged = {}
        for key in set(self.raw) | set(other.raw):
            merged[key] = op(self.raw.get(key, 0), other.raw.get(key, 0))
        return KubeResource(**merged)

    def __add__(self, other):
        return self._zip(other, lambda a, b: a + b)

    def __sub__(self, other):
        return self._zip(other, lambda a, b: a - b)

    def __mul__(self, multiplier):
        return KubeResource(**{key: value * multiplier for key, value in self.raw.items()})

    __rmul__ = __mul__

    def get(self, key, default=None):
        return self.raw.get(key, default)

    @property
    def possible(self):
        return all(value >= 0 for value in self.raw.values())

    def __str__(self):
        return str(self.raw)


_RESOURCES_MEMO = {}
_CREATED_BY_MEMO = {}
_TIME_MEMO = {}
_MISSING = object()


def _remember_time(text):
    value = utils.parse_time(text)  # raises like the reference for a malformed timestamp (not cached)
    if type(text) is str:
        if len(_TIME_MEMO) > (1 << 16):
            _TIME_MEMO.clear()
        _TIME_MEMO[text] = value
    return value


def _pod_resources(containers):
    """KubePod.resources (kube.py:41-49): pods:1 + the per-key sum over the containers' requests, accumulated as
    0.0 + v1 + v2 ... in container order (the order fixes the float64 result).  A cluster's pods come from a few
    templates, so the result is memoised on the literal request texts; pods with the same requests then SHARE one
    KubeResource (they are values: nothing in this package mutates one in place), which also lets the snapshot
    flattening convert each distinct vector once (snapshot.Dims).  Failures are not cached: they raise again."""
    try:
        if len(containers) == 1:  # the common case, without building a generator
            r = containers[0].get('resources')
            q = r.get('requests') if r else None
            key = tuple(q.items()) if q else ()
        else:
            key = []
            for c in containers:
                r = c.get('resources')
                q = r.get('requests') if r else None
                key.append(tuple(q.items()) if q else ())
            key = (len(containers), tuple(key))
        hit = _RESOURCES_MEMO.get(key)
    except TypeError:  # unhashable quantity objects: the plain route
        key, hit = None, None
    if hit is not None:
        return hit
    totals = {}
    for container in containers:
        for name, quantity in container.get('resources', {}).get('requests', {}).items():
            totals[name] = totals.get(name, 0.0) + utils.parse_SI(quantity)
    res = KubeResource(pods=1, **totals)
    if key is not None:
        if len(_RESOURCES_MEMO) > (1 << 16):
            _RESOURCES_MEMO.clear()
        _RESOURCES_MEMO[key] = res
    return res


class KubePod(object):
    _DRAIN_GRACE_PERIOD = datetime.timedelta(seconds=60 * 60)

    def __init__(self, pod):
        self.original = pod
        obj = pod.obj
        meta, spec, status = obj['metadata'], obj['spec'], obj['status']
        self.name = meta['name']
        self.namespace = meta['namespace']
        self.node_name = spec.get('nodeName')
        self.status = status['phase']
        self.uid = meta['uid']
        self.selectors = spec.get('nodeSelector', {})
        self.labels = labels = meta.get('labels', {})
        self.annotations = meta.get('annotations', {})
        self.owner = labels.get('owner', None)
        # (timestamps: memoised per text, utils.parse_time; the memo is consulted directly on this hot path)
        text = meta['creationTimestamp']
        cached = _TIME_MEMO.get(text) if type(text) is str else None
        self.creation_time = cached if cached is not None else _remember_time(text)
        text = status.get('startTime', _MISSING)
        if text is _MISSING:
            self.start_time = None
        else:
            cached = _TIME_MEMO.get(text) if type(text) is str else None
            self.start_time = cached if cached is not None else _remember_time(text)
        self.resources = _pod_resources(spec['containers'])

    def _created_by(self):
        """the created-by annotation, decoded (kube.py:51-58 decode it on every call).  The decoded value is a pure
        function of the text and only ever read, so it is memoised per distinct text (the pods of one controller
        share theirs); a malformed annotation is not cached and raises on every call, as upstream."""
        text = self.annotations.get('kubernetes.io/created-by', '{}')
        try:
            return _CREATED_BY_MEMO[text]
        except (KeyError, TypeError):
            pass
        decoded = json.loads(text)
        if type(text) is str:
            if len(_CREATED_BY_MEMO) > (1 << 16):
                _CREATED_BY_MEMO.clear()
            _CREATED_BY_MEMO[text] = decoded
        return decoded

    def is_mirrored(self):
        daemonset = self._created_by().get('reference', {}).get('kind') == 'DaemonSet'
        return daemonset or self.annotations.get('kubernetes.io/config.mirror')

    def is_replicated(self):
        return self._created_by()

    def is_critical(self):
        return utils.parse_bool_label(self.labels.get('openai/do-not-drain'))

    def is_in_drain_grace_period(self):
        """pods younger than an hour (or not started) are not drained (kube.py:63-68)."""
        if not self.start_time:
            return True
        return (utils.now(self.start_time.tzinfo) - self.start_time) < self._DRAIN_GRACE_PERIOD

    def is_drainable(self):
        return self.is_replicated() and not self.is_critical() and not self.is_in_drain_grace_period()

    def delete(self):
        logger.info('Deleting Pod %s/%s', self.namespace, self.name)
        return self.original.delete()

    def __hash__(self):
        return hash(self.uid)

    def __eq__(self, other):
        return self.uid == other.uid

    def __str__(self):
        return 'KubePod({namespace}, {name})'.format(namespace=self.namespace, name=self.name)

    __repr__ = __str__


_KUBEPOD_INIT = KubePod.__init__


def make_pods(raw_pods):
    """[KubePod(p) for p in raw_pods] (cluster.py:153 builds them with map).  With the C helper the plain pods (exact
    dicts, memoised timestamps and requests) are built in C attribute for attribute; everything else - and every
    pod when the helper is absent or KubePod.__init__ has been replaced - goes through the constructor."""
    if _hostfast is None or KubePod.__init__ is not _KUBEPOD_INIT:
        return list(map(KubePod, raw_pods))
    return _hostfast.make_pods(KubePod, raw_pods, _TIME_MEMO, _remember_time, _RESOURCES_MEMO, _pod_resources)


class KubeNode(object):
    def __init__(self, node):
        self.original = node
        self.pykube_node = node
        meta = node.obj['metadata']
        self.name = meta['name']
        self.index = int(self.name.split('-')[3])
        self.region, self.instance_type = self._get_instance_data()
        self.selectors = meta['labels']
        self.capacity = None  # set by Cluster.create_kube_node from the capacity table
        self.used_capacity = KubeResource()
        self.unschedulable = node.obj['spec'].get('unschedulable', False)
        self.creation_time = utils.parse_time(meta['creationTimestamp'])
        self.instance_index = utils.get_instance_index(node)

    def _get_instance_data(self):
        labels = self.original.obj['metadata']['labels']
        instance_type = labels.get('beta.kubernetes.io/instance-type')
        region = labels.get('failure-domain.beta.kubernetes.io/region')
        if instance_type and region:
            return (region, instance_type)
        return ('', None)

    # ---- adapter boundary: straight through to pykube (kube.py:124-167) -------------------------
    def drain(self, pods, notifier=None):
        for pod in pods:
            if pod.is_drainable():
                pod.delete()
        logger.info("drained %s", self)
        if notifier:
            notifier.notify_drained_node(self, pods)

    def _http_error(self):
        try:
            import pykube.exceptions
            return pykube.exceptions.HTTPError
        except ImportError:  # adapter not installed: nothing to catch
            return ()

    def uncordon(self):
        if not utils.parse_bool_label(self.selectors.get(_CORDON_LABEL)):
            logger.debug('uncordon %s ignored', self)
            return False
        try:
            self.original.reload()
            self.original.obj['spec']['unschedulable'] = False
            self.original.update()
            logger.info("uncordoned %s", self)
            return True
        except self._http_error() as ex:
            logger.info("uncordon failed %s %s", self, ex)
            return False

    def cordon(self):
        try:
            self.original.reload()
            self.original.obj['spec']['unschedulable'] = True
            self.original.obj['metadata']['labels'][_CORDON_LABEL] = 'true'
            self.original.update()
            logger.info("cordoned %s", self)
            return True
        except self._http_error() as ex:
            logger.info("cordon failed %s %s", self, ex)
            return False

    def delete(self):
        try:
            self.original.delete()
            logger.info("deleted %s", self)
            return True
        except self._http_error() as ex:
            logger.info("delete failed %s %s", self, ex)
            return False

    # ---- single-item forms of the hot-path arithmetic --------------------------------------------
    def count_pod(self, pod):
        assert isinstance(pod, KubePod)
        self.used_capacity += pod.resources

    def can_fit(self, resources):
        """capacity - (used + resources) possible (kube.py:173-176), evaluated on the GPU like the
        batched path (one pod x one node); Cluster.get_pending_pods never calls this per pair."""
        assert isinstance(resources, KubeResource)
        from . import snapshot
        return snapshot.single_can_fit(self, resources)

    def is_match(self, pod):
        return all(self.selectors.get(label) == value for label, value in pod.selectors.items())

    def __hash__(self):
        return hash(self.name)

    def __eq__(self, other):
        return self.name == other.name

    def __str__(self):
        return "{}".format(self.name)
