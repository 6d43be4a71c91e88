#!/bin/bash
# round-2 run B: compact pooled kernel (one scan / one drain site, tight direction bound, global-memory scan for far pixels)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2b_gpu_tests.log
for v in "" r1 abl1 abl2; do
  VXL_LIB=${v:+$PWD/tools/exp/variants/$v.so} python tools/exp/time_ambient.py 3 7
done > gpurun_out/r2b_split.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ambient -c 1 -f -o gpurun_out/prof_r2b_ao python tools/exp/prof_ao.py ao 1 > gpurun_out/r2b_ncu_ao.log 2>&1
cat gpurun_out/r2b_gpu_tests.log; grep -v "^ \|Traceback" gpurun_out/r2b_split.log | cut -c1-330
