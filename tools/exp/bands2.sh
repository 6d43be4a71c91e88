#!/bin/bash
# e2e (vxl_lighting_host / _packed) against the number of row bands
mkdir -p gpurun_out
for nb in 2 3 4 6 8 12 16 24; do
  VXL_HOST_BANDS=$nb python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('NB=$nb', 'resident', round(d['ms_per_step'],3), 'packed', round(e['ms_per_step'],3), 'float', round(e['float_planes']['ms_per_step'],3))" | tee -a gpurun_out/bands2.log
done
