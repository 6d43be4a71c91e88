#!/bin/bash
# round-2 run A: parity of the pooled AO resolve, A/B timings of k_ambient variants, read-bandwidth probe, source-level profiles
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2a_gpu_tests.log
for v in "" r1 q0b2 nopair abl1 abl2; do
  VXL_LIB=${v:+$PWD/tools/exp/variants/$v.so} python tools/exp/time_ambient.py 3 7
done > gpurun_out/r2a_split.log 2>&1
python - > gpurun_out/r2a_bw.log 2>&1 <<'PY'
from voxelengine_b200 import engine as E
c = E.Context(0)
for mb in (8, 16, 32, 48, 64, 96, 128, 256, 1024, 4096):
    print(mb, "MiB", round(c.read_bandwidth(mb << 20, max(2, 4096 // mb)), 1), "GB/s", flush=True)
PY
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench tools/ubench.cu 2>/dev/null && /tmp/ubench > gpurun_out/r2a_ubench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ambient -c 1 -f -o gpurun_out/prof_r2a_ao python tools/exp/prof_ao.py ao 1 > gpurun_out/r2a_ncu_ao.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ambient -c 1 -f -o gpurun_out/prof_r2a_sun python tools/exp/prof_ao.py sun 1 > gpurun_out/r2a_ncu_sun.log 2>&1
cat gpurun_out/r2a_gpu_tests.log gpurun_out/r2a_split.log gpurun_out/r2a_bw.log
