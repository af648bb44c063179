"""Downward adapter boundary (cloud / template I/O) -- NOT re-implemented here.

The hot path ends where the reference hands `new_pool_sizes` to `scale_pools` and node
actions to kube.py's mutators.  Everything below that line (Azure login, ARM template
download / unrolling, deployments, VM deletion: reference autoscaler/azure_api.py,
template_processing.py) is cloud I/O that a deployment keeps from the reference package.
This module resolves those functions lazily from a module named by $ACSFIT_ADAPTER_MODULE, from this package's
own template_processing.py (the ARM template unrolling), or from the reference's modules when
they are importable (`autoscaler.azure_api`), and raises a clear error otherwise.  With --dry-run none of them
is ever called (engine_scaler.py:79-88, :154-178).
"""
import importlib
import os

_NAMES = {
    "login": "azure_api", "download_template": "azure_api", "download_parameters": "azure_api",
    "create_deployment": "azure_api", "delete_resources_for_node": "azure_api",
    "prepare_template_for_scale_out": "template_processing", "delete_master_vm_extension": "template_processing",
}
_overrides = {}


def register(name, fn):
    """install an adapter function explicitly (used by tests and by embedding applications)."""
    if name not in _NAMES:
        raise KeyError(name)
    _overrides[name] = fn


def resolve(name):
    if name in _overrides:
        return _overrides[name]
    candidates = []
    if os.environ.get("ACSFIT_ADAPTER_MODULE"):
        candidates.append(os.environ["ACSFIT_ADAPTER_MODULE"])
    if _NAMES[name] == "template_processing":  # the template unrolling is part of this package now
        candidates.append(__package__ + ".template_processing")
    candidates.append("autoscaler." + _NAMES[name])
    for mod in candidates:
        try:
            return getattr(importlib.import_module(mod), name)
        except (ImportError, AttributeError):
            continue
    raise RuntimeError("adapter function %r is not available: install the reference's autoscaler.%s "
                       "(cloud I/O is outside this package) or register one with adapters.register()"
                       % (name, _NAMES[name]))


def call(name, *args, **kwargs):
    return resolve(name)(*args, **kwargs)
