#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_full_size.py -m gpu -x -q -k "lighting_host or packed" 2>&1 | tail -3
for nb in "" 8 12 16; do
  VXL_HOST_BANDS=$nb python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('NB=$nb', 'resident', round(d['ms_per_step'],3), 'packed', round(e['ms_per_step'],3), 'float', round(e['float_planes']['ms_per_step'],3))" | tee -a gpurun_out/bands3.log
done
for c in 2 4; do python bench.py --config $c --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('cfg$c', 'resident', round(d['ms_per_step'],3), 'packed', round(e['ms_per_step'],3), 'float', round(e['float_planes']['ms_per_step'],3))" | tee -a gpurun_out/bands3.log; done
