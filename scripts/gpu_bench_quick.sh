# usage: bash scripts/gpu_bench_quick.sh <tag> [pytest -k expr]
set +e
TAG=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x ${2:+-k "$2"} > gpurun_out/${TAG}_pytest.log 2>&1
echo "exit $?" >> gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --cpu-budget 2 > gpurun_out/${TAG}_bench.log 2>&1
echo "exit $?" >> gpurun_out/${TAG}_bench.log
tail -n 6 gpurun_out/${TAG}_pytest.log
python - <<PY
import json
for l in open('gpurun_out/${TAG}_bench.log'):
    if l.startswith('{'):
        d = json.loads(l)
        print('value %.0f frames/s  %.3f ms/step  e2e %.0f (%.3f ms)' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']))
        for k in d['kernels']:
            print('  %-26s %.4f ms  %s' % (k['kernel'], k['ms_per_step'], ('%.1f%% of %s' % (100 * k['frac'], k['bound'])) if 'frac' in k else ''))
    elif 'rror' in l or 'exit' in l:
        print(l.strip())
PY
