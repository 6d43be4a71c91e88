python -m pytest tests/test_taa_reflection.py tests/test_resolve.py -m gpu -x -q 2>&1 | tail -2; python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e 2>gpurun_out/exp_bench.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['light_buffer_resolve']['ms']); print(d['post_passes']); print(d['geometry_pass']['ms'])"
