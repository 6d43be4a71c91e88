#!/bin/bash
# round-2 run J: stragglers? pixels by mode dropped / resolve dropped, 32x16 regions
mkdir -p gpurun_out
for v in "" skip3 skip2 abl1 abl1s2 cnt2; do
  VXL_LIB=${v:+$PWD/tools/exp/variants/$v.so} python tools/exp/time_ambient.py 3 5
done > gpurun_out/r2j_split.log 2>&1
grep -v "^ \|Traceback" gpurun_out/r2j_split.log | cut -c1-1200
