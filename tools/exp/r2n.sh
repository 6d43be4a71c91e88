#!/bin/bash
# round-2 run N: closed-form group tests (Sparse march), one static item per warp
mkdir -p gpurun_out
true
cat gpurun_out/r2n_gpu_tests.log
for v in "" cf1 one1; do
  VXL_EXP_NAO=1 VXL_LIB=${v:+$PWD/tools/exp/variants/$v.so} python tools/exp/time_ambient.py 3 5
done > gpurun_out/r2n_split.log 2>&1
grep -v "^ \|Traceback" gpurun_out/r2n_split.log | cut -c1-1200
for v in "" cf1 one1; do
  VXL_LIB=${v:+$PWD/tools/exp/variants/$v.so} python bench.py --steps 10 --warmup 3 > gpurun_out/r2n_bench_${v:-default}.json 2> gpurun_out/r2n_bench_${v:-default}.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2n_bench_${v:-default}.json").read().strip().splitlines()[-1])
    print("${v:-default}", "value", d.get("value"), "ms", d.get("ms_per_step"), "kernels", d.get("roofline", {}).get("all_kernels_ms"), "frac", d.get("roofline", {}).get("frac"), "e2e", d.get("e2e", {}).get("ms_per_step"), d.get("e2e", {}).get("float_planes", {}).get("ms_per_step"))
except Exception as e:
    print("bench ${v:-default} failed", e); print(open("gpurun_out/r2n_bench_${v:-default}.err").read()[-2000:])
PY
done
