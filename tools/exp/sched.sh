#!/bin/bash
# e2e (packed) against the band schedule of vxl_lighting_host_packed
mkdir -p gpurun_out
for sc in "" "1,5,5,4,1" "1,3,4,4,3,1" "1,2,3,3,3,2,1,1" "1,2,4,4,4,4,4,4,2,1" "1,6,6,2,1"; do
  if [ -z "$sc" ]; then unset VXL_HOST_SCHED; else export VXL_HOST_SCHED=$sc; fi
  python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('sched=[$sc]', 'resident', round(d['ms_per_step'],3), 'packed', round(e['ms_per_step'],3), 'float', round(e['float_planes']['ms_per_step'],3))" | tee -a gpurun_out/sched.log
done
