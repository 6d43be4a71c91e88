#!/bin/bash
# multi-GPU: fused gather with the flag fence (vxl_group), config 3
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2s_bench_n$N.json 2> gpurun_out/r2s_bench_n$N.err
tail -c 600 gpurun_out/r2s_bench_n$N.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2s_bench_n$N.json").read().strip().splitlines()[-1])
print("N=$N value", d.get("value"), "ms", d.get("ms_per_step"), "kernels", d.get("roofline", {}).get("all_kernels_ms"), "e2e", d.get("e2e", {}).get("ms_per_step"), d.get("e2e", {}).get("value"), d.get("detail", {}).get("parallelism", "")[:120])
PY
