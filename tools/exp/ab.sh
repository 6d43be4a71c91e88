# A/B two builds of libvxl.so on the same box: default vs tools/exp/libvxl_p0.so
for lib in "" tools/exp/libvxl_p0.so; do
  echo "== lib=${lib:-default}"
  VXL_LIB=${lib:+$PWD/$lib} python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), d['roofline']['all_kernels_ms'], d['config']['probes_that_read_the_volume'])"
done
