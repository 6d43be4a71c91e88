N=${1:-2}
python bench.py --config 4 --steps 10 --warmup 3 > gpurun_out/r1g_bench_cfg4.json 2> gpurun_out/r1g_bench_cfg4.err
for c in 3 5; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --config $c --steps 10 --warmup 3 > gpurun_out/r1g_bench_cfg${c}_n$N.json 2> gpurun_out/r1g_bench_cfg${c}_n$N.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/r1g_bench_ref_n$N.json 2> gpurun_out/r1g_bench_ref_n$N.err
