#!/bin/bash
# round-2 run D: where k_ambient's AO time goes -- pixels by scan mode, fixed cost vs per-ray slope
mkdir -p gpurun_out
for v in "" skip3 skip2 fixed abl2 cnt2 cnt3; do
  VXL_EXP_NAO=1,2,4,8 VXL_LIB=${v:+$PWD/tools/exp/variants/$v.so} python tools/exp/time_ambient.py 3 5
done > gpurun_out/r2d_split.log 2>&1
grep -v "^ \|Traceback" gpurun_out/r2d_split.log | cut -c1-1200
