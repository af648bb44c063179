#!/bin/bash
# developer helper: run bench.py and print the key numbers of its JSON line
python bench.py "$@" 2>&1 | python -c '
import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        b = json.loads(l)
        print("value %.1f G/s  ms/step %.3f  e2e ms %.3f  kernel ms %.3f  launches %d  frac %.3f" % (
            b["value"] / 1e9, b["ms_per_step"], b["e2e"]["ms_per_step"], b["roofline"]["kernel_ms_per_step"],
            b["gpu_launches"], b["roofline"]["frac"]))
    else:
        print(l.rstrip()[:300])
'
