#!/bin/bash
# A/B of library variants on the rows either side of the path (config 3): TAA, resolves, geometry pass
for v in "$@"; do
  lib=""; [ "$v" != default ] && lib=$PWD/tools/exp/variants/$v.so
  VXL_LIB=$lib python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); p=d['post_passes']; print('$v', 'frame', round(d['ms_per_step'],3), 'taa', round(p['taa_ms'],4), 'refl_colour', round(p['reflection_colour_ms'],4), 'resolve', round(d['light_buffer_resolve']['ms'],4), 'geometry', round(d['geometry_pass']['ms'],4))"
done
