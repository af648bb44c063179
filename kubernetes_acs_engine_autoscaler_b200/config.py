"""Process-wide settings taken from the environment -- the same two variables the reference reads
(autoscaler/config.py:4-6), so an existing deployment's manifest keeps working.

CAPACITY_DATA         path of the instance-type table: a JSON object {type: {resource: amount}} whose key
                      ORDER is the pools' cost order (reference capacity.py:34-36).
CAPACITY_CPU_RESERVE  CPU (cores, float) held back on every instance type for system pods.
"""
import os


def _from_env(name, default, convert=str):
    raw = os.environ.get(name)
    return default if raw is None else convert(raw)


class Config(object):
    pass


def _default_capacity_data():
    """the reference's default is the cwd-relative 'data/capacity.json' (config.py:5); when that file is not
    there, the copy of the same constant table shipped at the root of this repository is used."""
    if os.path.exists('data/capacity.json'):
        return 'data/capacity.json'
    return os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'data', 'capacity.json')


Config.CAPACITY_DATA = _from_env('CAPACITY_DATA', _default_capacity_data())
Config.CAPACITY_CPU_RESERVE = _from_env('CAPACITY_CPU_RESERVE', 0.0, float)
