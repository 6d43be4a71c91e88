# A/B of vxl_lighting (three pass kernels on concurrent streams) against one pass after the other.  usage: conc_ab.sh <N>
N=$1
for v in 0 1; do
if [ "$N" = "1" ]; then
VXL_CONCURRENT=$v python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e 2>gpurun_out/conc_$v.err > gpurun_out/conc_${v}_n1.json
else
VXL_CONCURRENT=$v python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --config 3 --steps 10 --warmup 3 --no-e2e > gpurun_out/conc_${v}_n$N.json 2> gpurun_out/conc_$v.err
fi
tail -2 gpurun_out/conc_$v.err | grep -v "OMP\|\*\*\*\|NCCL version"
python -c "
import json; d=json.load(open('gpurun_out/conc_${v}_n$N.json')); print('concurrent=$v', d['n_gpus'], round(d['ms_per_step'],3), round(d['value']), d['config'].get('gather'), d['gpu_launches'])"
done
