python -m pytest tests -m gpu -x -q 2>&1 | tail -3; python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e 2>gpurun_out/exp_bench.err | tee gpurun_out/exp_bench.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['all_kernels_ms'], d['roofline']['frac']); print(d['light_buffer_resolve']); print(d['post_passes']); print(d['geometry_pass'])"
tail -3 gpurun_out/exp_bench.err
