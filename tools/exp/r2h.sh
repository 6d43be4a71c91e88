#!/bin/bash
# round-2 run H (region blocks): TMA tile + tiles centred on the ray boxes: parity, k_ambient split, pixels by scan mode, per-kernel frame times
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2h_gpu_tests.log
cat gpurun_out/r2h_gpu_tests.log
for v in "" cnt2 cnt3; do
  VXL_EXP_NAO=1 VXL_LIB=${v:+$PWD/tools/exp/variants/$v.so} python tools/exp/time_ambient.py 3 5
done > gpurun_out/r2h_split.log 2>&1
grep -v "^ \|Traceback" gpurun_out/r2h_split.log | cut -c1-1200
for v in "" pb2; do
  VXL_LIB=${v:+$PWD/tools/exp/variants/$v.so} python bench.py --steps 10 --warmup 3 > gpurun_out/r2h_bench_${v:-default}.json 2> gpurun_out/r2h_bench_${v:-default}.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2h_bench_${v:-default}.json").read().strip().splitlines()[-1])
    print("${v:-default}", "value", d.get("value"), "ms", d.get("ms_per_step"), "kernels", d.get("roofline", {}).get("all_kernels_ms"), "frac", d.get("roofline", {}).get("frac"), "e2e", d.get("e2e", {}).get("ms_per_step"), d.get("e2e", {}).get("float_planes", {}).get("ms_per_step"))
except Exception as e:
    print("bench ${v:-default} failed", e); print(open("gpurun_out/r2h_bench_${v:-default}.err").read()[-2000:])
PY
done
