#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_ambient -c 1 -f -o gpurun_out/prof_r2g_ao python tools/exp/prof_ao.py ao 1 > gpurun_out/r2g_ncu_ao.log 2>&1
tail -3 gpurun_out/r2g_ncu_ao.log
