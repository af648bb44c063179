#!/bin/bash
# quick GPU check of a resolver change: parity subset, then the default bench line (c2 headline + c3 sub-record)
tag=${1:-q}
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_mid_golden_gpu.py -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
c3=d.get("configs",{}).get("c3",{})
print("c2 ms", d["ms_per_step"], "e2e", d.get("e2e",{}).get("ms_per_step"), "spread", d.get("ms_per_step_spread"))
print("c3 ms", c3.get("ms_per_step"), {k:v for k,v in c3.items() if "ms" in k or k=="passes"})
print("py surface", d.get("e2e",{}).get("python_surface"))
PY
