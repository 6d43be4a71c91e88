#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_full_size.py -m gpu -x -q --durations=8 2>&1 | tail -25 > gpurun_out/r2q_gpu_tests.log
cat gpurun_out/r2q_gpu_tests.log
