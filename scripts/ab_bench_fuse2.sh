for i in 1 2; do
for f in "--no-fuse2" ""; do
python bench.py --steps 30 --warmup 5 --latency-frames 0 --cpu-kind port --cpu-budget 1 $f 2>/dev/null | python -c "import json,sys; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$f', round(l['value']), round(l['ms_per_step'],4), round(l['e2e']['value']), l['clocks']['sm_mhz'], l['clocks']['reasons'], [(k['kernel'][:8], round(k['ms_per_step'],3)) for k in l['kernels']])"
done; done
