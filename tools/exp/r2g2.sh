#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'k_gbuffer_models' -c 1 \
    -f -o gpurun_out/prof_r2g2 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2g2_ncu.log 2>&1
tail -2 gpurun_out/r2g2_ncu.log
