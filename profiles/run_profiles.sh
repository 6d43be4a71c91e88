#!/bin/bash
# Run on the GPU box (under gpurun): launch list + one full ncu capture of the pass kernels.
# usage: profiles/run_profiles.sh <tag> [config]
TAG=${1:-r1}
CFG=${2:-3}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --config $CFG --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_ambient|k_local_lights|k_reflection|k_resolve' -s 9 -c 5 \
    -f -o gpurun_out/prof_${TAG} python bench.py --config $CFG --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out
