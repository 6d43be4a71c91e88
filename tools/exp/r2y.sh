#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r2y_gpu_tests.log
N=2
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2y_bench_n$N.json 2> gpurun_out/r2y_bench_n$N.err
tail -c 1500 gpurun_out/r2y_bench_n$N.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2y_bench_n$N.json").read().strip().splitlines()[-1])
print("N=$N value", d.get("value"), "ms", d.get("ms_per_step"), "kernels", d.get("roofline", {}).get("all_kernels_ms"), "e2e", d.get("e2e"))
PY
python bench.py > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2y_bench.json").read().strip().splitlines()[-1])
print("value", d.get("value"), "ms", d.get("ms_per_step"), "kernels", d.get("roofline", {}).get("all_kernels_ms"), "frac", d.get("roofline", {}).get("frac"), "e2e", d.get("e2e", {}).get("ms_per_step"), d.get("e2e", {}).get("float_planes", {}).get("ms_per_step"))
print("post", d.get("post_passes"), "resolve", d.get("light_buffer_resolve"))
PY
