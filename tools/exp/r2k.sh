#!/bin/bash
# round-2 run K: negative coordinates on the scan path, warp-uniform scan source
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2k_gpu_tests.log
cat gpurun_out/r2k_gpu_tests.log
for v in "" noprom cnt2 cnt3 skip2; do
  VXL_LIB=${v:+$PWD/tools/exp/variants/$v.so} python tools/exp/time_ambient.py 3 5
done > gpurun_out/r2k_split.log 2>&1
grep -v "^ \|Traceback" gpurun_out/r2k_split.log | cut -c1-1200
