#!/bin/bash
# full ncu capture of the f-row kernels (TAA, resolves, geometry pass) on config 3
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'k_light_taa|k_resolve|k_gbuffer_models' -c 7 \
    -f -o gpurun_out/prof_r2x python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2x_ncu.log 2>&1
tail -3 gpurun_out/r2x_ncu.log
ls -la gpurun_out/prof_r2x.ncu-rep
