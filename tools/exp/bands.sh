for nb in 4 6 8 12 16; do
  VXL_HOST_BANDS=$nb python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bands', $nb, 'e2e ms', d['e2e']['ms_per_step'], 'kernel ms', d['ms_per_step'])"
done
