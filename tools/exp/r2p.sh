#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2p_gpu_tests.log
cat gpurun_out/r2p_gpu_tests.log
