#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2c_gpu_tests.log
for v in "" r1; do
  VXL_LIB=${v:+$PWD/tools/exp/variants/$v.so} python tools/exp/time_ambient.py 3 7
done > gpurun_out/r2c_split.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ambient -c 1 -f -o gpurun_out/prof_r2c_ao python tools/exp/prof_ao.py ao 1 > gpurun_out/r2c_ncu_ao.log 2>&1
cat gpurun_out/r2c_gpu_tests.log; grep -v "^ \|Traceback" gpurun_out/r2c_split.log | cut -c1-330
