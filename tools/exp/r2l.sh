#!/bin/bash
# round-2 run L: several tile placements per block (k_ambient)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2l_gpu_tests.log
cat gpurun_out/r2l_gpu_tests.log
for v in "" rnd1 rnd2 cnt2; do
  VXL_EXP_NAO=1 VXL_LIB=${v:+$PWD/tools/exp/variants/$v.so} python tools/exp/time_ambient.py 3 5
done > gpurun_out/r2l_split.log 2>&1
grep -v "^ \|Traceback" gpurun_out/r2l_split.log | cut -c1-1200
