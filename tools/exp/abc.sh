#!/bin/bash
# A/B library variants on a given config: tools/exp/abc.sh <config> name1 name2 ...
C=$1; shift
for v in "$@"; do
  lib=""; [ "$v" != default ] && lib=$PWD/tools/exp/variants/$v.so
  for rep in 1 2; do
  VXL_LIB=$lib python bench.py --config $C --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg$C $v', round(d['ms_per_step'],4), {k: round(x,4) for k,x in d['roofline']['all_kernels_ms'].items()})"
  done
done
