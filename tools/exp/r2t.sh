#!/bin/bash
mkdir -p gpurun_out
CUDA_LAUNCH_BLOCKING=1 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pass_reflection or pass_point" 2>&1 | tail -30 > gpurun_out/r2t.log
cut -c1-300 gpurun_out/r2t.log | tail -30
