N=$1
for g in nccl fused; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --config 3 --steps 10 --warmup 3 --no-e2e --gather $g > gpurun_out/ab_${g}_n$N.json 2> gpurun_out/ab_${g}_n$N.err
tail -3 gpurun_out/ab_${g}_n$N.err
python -c "
import json; d=json.load(open('gpurun_out/ab_${g}_n$N.json')); print('$g', d['n_gpus'], round(d['ms_per_step'],3), round(d['value']), d['config']['gather'], d['roofline']['all_kernels_ms'])"
done
