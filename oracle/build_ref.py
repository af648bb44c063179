#!/usr/bin/env python3
"""TEST INFRASTRUCTURE -- recipe for oracle/_ref: the UNMODIFIED reference, staged so that it travels.

The reference is pure Python (there is nothing to compile), and /root/reference does not exist on the GPU box.
This recipe copies the files of the hot path verbatim from where they lie (autoscaler/*.py, data/capacity.json)
into the git-ignored oracle/_ref/ (listed in .gitignore, NOT in .gpurunignore, so it ships with the snapshot
like a built .so).  bench.py's cpu_baseline / --impl reference legs import it from there under
oracle/ref_shim.py to time the reference's own CPython loop next to the GPU numbers; nothing in the product
imports it, and no reference source is committed.  Run by __graft_entry__.build() when /root/reference exists.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("ACSFIT_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")


def build(force=False):
    """returns the staged root, or None when the reference tree is not here (the GPU box: use what shipped)."""
    if not os.path.isdir(os.path.join(SRC, "autoscaler")):
        return DST if os.path.isdir(os.path.join(DST, "autoscaler")) else None
    marker = os.path.join(DST, "autoscaler", "cluster.py")
    if os.path.exists(marker) and not force:
        return DST
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(os.path.join(DST, "data"))
    shutil.copytree(os.path.join(SRC, "autoscaler"), os.path.join(DST, "autoscaler"),
                    ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    shutil.copy2(os.path.join(SRC, "data", "capacity.json"), os.path.join(DST, "data", "capacity.json"))
    return DST


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
