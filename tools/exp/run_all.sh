# one-box capture: GPU tests, bench lines of all configs + the reference arm, launch list + full ncu capture.  usage: run_all.sh <tag>
TAG=${1:-r1}
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/gpu_tests.log
python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
for c in 1 2 4 5; do python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_cfg$c.json 2> gpurun_out/${TAG}_bench_cfg$c.err; done
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2> gpurun_out/${TAG}_bench_ref.err
bash profiles/run_profiles.sh ${TAG}
