# usage: bash scripts/gpu_check.sh <tag> [pytest -k expr]
set +e
TAG=${1:-chk}
mkdir -p gpurun_out
echo "=== pytest" > gpurun_out/${TAG}_pytest.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x ${2:+-k "$2"} >> gpurun_out/${TAG}_pytest.log 2>&1
echo "exit $?" >> gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1
echo "exit $?" >> gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.log 2>&1
echo "exit $?" >> gpurun_out/${TAG}_bench.log
tail -n 12 gpurun_out/${TAG}_pytest.log; tail -n 3 gpurun_out/${TAG}_smoke.log; tail -n 3 gpurun_out/${TAG}_bench.log
