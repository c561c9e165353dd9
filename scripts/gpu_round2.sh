set +e
mkdir -p gpurun_out
echo "=== pytest all" > gpurun_out/r2_pytest.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 >> gpurun_out/r2_pytest.log 2>&1
echo "exit $?" >> gpurun_out/r2_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1
echo "exit $?" >> gpurun_out/r2_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench.log 2>&1
echo "exit $?" >> gpurun_out/r2_bench.log
timeout 600 python bench.py --steps 5 --warmup 3 --gemm-impl 3 --cpu-budget 2 > gpurun_out/r2_bench_v1gemm.log 2>&1
echo "exit $?" >> gpurun_out/r2_bench_v1gemm.log
tail -n 8 gpurun_out/r2_pytest.log gpurun_out/r2_smoke.log
