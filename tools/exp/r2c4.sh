#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r2c4.csv \
    python bench.py --config 4 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2c4.log 2>&1
tail -2 gpurun_out/r2c4.log | cut -c1-300
