#!/usr/bin/env python3
"""Writes data/capacity.json (the instance-type table Config.CAPACITY_DATA defaults to) from the values
recorded in tests/golden/capacity_table.json, which oracle/make_golden.py took from the reference's
capacity.RESOURCE_SPEC (reference data/capacity.json:1-129, CAPACITY_CPU_RESERVE = 0).  It is a constant
table of Azure VM sizes; key ORDER is the pools' cost order (capacity.py:34-36) and the mistyped " pods" key
of the last row is kept because it changes results."""
import json
import os
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    with open(os.path.join(ROOT, "tests", "golden", "capacity_table.json")) as f:
        rows = json.load(f, object_pairs_hook=OrderedDict)["rows"]
    table = OrderedDict()
    for name, amounts in rows:
        table[name] = OrderedDict((k, float.fromhex(v)) for k, v in amounts.items())
    out = os.path.join(ROOT, "data", "capacity.json")
    with open(out, "w") as f:
        f.write("{\n")
        items = list(table.items())
        for i, (name, amounts) in enumerate(items):
            f.write("  %s: %s%s\n" % (json.dumps(name), json.dumps(amounts), "," if i + 1 < len(items) else ""))
        f.write("}\n")
    print(out, len(table), "instance types")


if __name__ == "__main__":
    main()
