#!/bin/bash
# A/B differently compiled copies of libvxl.so on one box: tools/exp/abv.sh <tag> name1 name2 ... ("default" = the in-tree library)
T=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  lib=""; [ "$v" != default ] && lib=$PWD/tools/exp/variants/$v.so
  for rep in 1 2; do
  VXL_LIB=$lib python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', round(d['ms_per_step'],3), {k: round(x,3) for k,x in d['roofline']['all_kernels_ms'].items()})" | tee -a gpurun_out/${T}_ab.log
  done
done
