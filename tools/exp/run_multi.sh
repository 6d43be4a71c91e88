# multi-GPU bench lines (one box, N ranks): configs 3 and 5, plus the reference arm under torchrun.  usage: run_multi.sh <N> <tag>
N=${1:-2}
TAG=${2:-r1}
for c in 3 5; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --config $c --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_cfg${c}_n$N.json 2> gpurun_out/${TAG}_bench_cfg${c}_n$N.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref_n$N.json 2> gpurun_out/${TAG}_bench_ref_n$N.err
