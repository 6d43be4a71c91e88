#!/bin/bash
# full GPU suite + default bench on the current tree
T=${1:-r2v}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${T}_gpu_tests.log
cat gpurun_out/${T}_gpu_tests.log
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${T}_bench.json").read().strip().splitlines()[-1])
print("value", d.get("value"), "ms", d.get("ms_per_step"), "kernels", d.get("roofline", {}).get("all_kernels_ms"), "frac", d.get("roofline", {}).get("frac"), "e2e", d.get("e2e", {}).get("ms_per_step"), d.get("e2e", {}).get("float_planes", {}).get("ms_per_step"))
PY
