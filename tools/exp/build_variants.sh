#!/bin/bash
# Experiment helper: differently compiled copies of libvxl.so for A/B runs on one GPU box (loaded through VXL_LIB).
# usage: tools/exp/build_variants.sh name1 "flags1" name2 "flags2" ...
mkdir -p tools/exp/variants
while [ $# -ge 2 ]; do
  ( VXL_LIB_OUT=$PWD/tools/exp/variants/$1.so VXL_NVCC_EXTRA="$2" python -m voxelengine_b200.build --force > /dev/null && echo "built $1: $2" ) &
  shift 2
done
wait
