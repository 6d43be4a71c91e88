#!/bin/bash
# round-2 profiles: launch list + full ncu capture of the pass kernels (config 3), DRAM traffic captures for configs 2 and 5
mkdir -p gpurun_out
profiles/run_profiles.sh r2 3 > /dev/null 2>&1
for C in 2 5; do
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:'k_ambient|k_local_lights|k_reflection' -s 9 -c 3 \
      --csv --log-file gpurun_out/traffic_r2_cfg$C.csv python bench.py --config $C --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/traffic_r2_cfg$C.log 2>&1
done
ls -la gpurun_out | grep r2 | tail -12
