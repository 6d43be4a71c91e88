#!/bin/bash
# round-2 run I: region size sweep
mkdir -p gpurun_out
for v in "" r32x32 r64x16 r32x16 r128x32; do
  VXL_EXP_NAO=1 VXL_LIB=${v:+$PWD/tools/exp/variants/$v.so} python tools/exp/time_ambient.py 3 5
done > gpurun_out/r2i_split.log 2>&1
grep -v "^ \|Traceback" gpurun_out/r2i_split.log | cut -c1-1200
