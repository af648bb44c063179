"""Instance-type capacity table (mirror of reference autoscaler/capacity.py).

RESOURCE_SPEC[type] is the KubeResource a fresh node of that type offers (file value, cpu minus
CAPACITY_CPU_RESERVE); the JSON key order is the pools' cost order.  Loaded at import from
Config.CAPACITY_DATA exactly like the reference (capacity.py:12-18) -- quirks of the data file
(e.g. a mistyped " pods" key) are preserved because they change results.
"""
import json
from collections import OrderedDict

from .config import Config
from .kube import KubeResource

DEFAULT_TYPE_SELECTOR_KEY = 'beta.kubernetes.io/instance-type'


def load(path=None, cpu_reserve=None):
    """(re)load the table; returns (data, RESOURCE_SPEC)."""
    global data, RESOURCE_SPEC
    path = Config.CAPACITY_DATA if path is None else path
    reserve = Config.CAPACITY_CPU_RESERVE if cpu_reserve is None else cpu_reserve
    with open(path, 'r') as f:
        table = json.loads(f.read(), object_pairs_hook=OrderedDict)
    spec = {}
    for instance_type, amounts in table.items():
        amounts['cpu'] -= reserve
        spec[instance_type] = KubeResource(**amounts)
    data, RESOURCE_SPEC = table, spec
    return data, RESOURCE_SPEC


data, RESOURCE_SPEC = OrderedDict(), {}
load()  # like the reference (capacity.py:12-18): a missing table is an import-time error, not a KeyError later


def get_capacity_for_instance_type(instance_type):
    return RESOURCE_SPEC[instance_type]


def is_possible(pod, agent_pools):
    """whether the pod fits an empty instance of at least one pool (capacity.py:24-32); single-pod
    form of the K0 kernel (Cluster.get_pods_to_schedule batches all pods into one launch)."""
    from . import snapshot
    return bool(snapshot.feasible_pods([pod], agent_pools)[0])


def order_by_cost_asc(agent_pools):
    keys = list(data.keys())
    return sorted(agent_pools, key=lambda pool: keys.index(pool.instance_type))
