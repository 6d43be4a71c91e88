#!/bin/bash
mkdir -p gpurun_out
VXL_EXP_NAO1=1 ncu --set full --clock-control none --import-source on -k regex:k_ambient -c 1 -f -o gpurun_out/prof_r2m_ao1 python tools/exp/prof_ao.py ao 1 > gpurun_out/r2m_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ambient -c 1 -f -o gpurun_out/prof_r2m_ao python tools/exp/prof_ao.py ao 1 >> gpurun_out/r2m_ncu.log 2>&1
tail -2 gpurun_out/r2m_ncu.log
