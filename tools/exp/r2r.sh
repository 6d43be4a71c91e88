#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lighting_host" 2>&1 | tail -15 > gpurun_out/r2r_gpu_tests.log
cat gpurun_out/r2r_gpu_tests.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2r_bench.json").read().strip().splitlines()[-1])
print("value", d.get("value"), "ms", d.get("ms_per_step"), "kernels", d.get("roofline", {}).get("all_kernels_ms"), "frac", d.get("roofline", {}).get("frac"), "e2e", d.get("e2e", {}).get("ms_per_step"), d.get("e2e", {}).get("float_planes", {}).get("ms_per_step"))
PY
