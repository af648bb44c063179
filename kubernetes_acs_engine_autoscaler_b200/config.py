"""Environment-driven settings (same variables as reference autoscaler/config.py:4-6)."""
import os


class Config(object):
    # path of the instance-type table (JSON object: type -> {resource: amount}); its key ORDER is
    # the pools' cost order (reference capacity.py:34-36)
    CAPACITY_DATA = os.environ.get('CAPACITY_DATA', 'data/capacity.json')
    # CPU held back on every instance type for system pods
    CAPACITY_CPU_RESERVE = float(os.environ.get('CAPACITY_CPU_RESERVE', 0.0))
