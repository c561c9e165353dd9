set +e
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
ls /root/reference > gpurun_out/refprobe.txt 2>&1
echo "=== pytest SIMT gemm" > gpurun_out/r1_pytest_simt.log
B200POSE_GEMM_IMPL=1 timeout 900 python -m pytest tests -m gpu -q --timeout 300 -k "not test_linear_kernels and not full_size" >> gpurun_out/r1_pytest_simt.log 2>&1
echo "exit $?" >> gpurun_out/r1_pytest_simt.log
echo "=== pytest linear kernels" > gpurun_out/r1_pytest_linear.log
timeout 600 python -m pytest tests -m gpu -q --timeout 200 -k "test_linear_kernels" >> gpurun_out/r1_pytest_linear.log 2>&1
echo "exit $?" >> gpurun_out/r1_pytest_linear.log
echo "=== pytest all (tcgen05)" > gpurun_out/r1_pytest_tc.log
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 >> gpurun_out/r1_pytest_tc.log 2>&1
echo "exit $?" >> gpurun_out/r1_pytest_tc.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1_smoke.log 2>&1
echo "exit $?" >> gpurun_out/r1_smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r1_bench.log 2>&1
echo "exit $?" >> gpurun_out/r1_bench.log
timeout 600 python bench.py --steps 3 --warmup 3 --gemm-impl 1 > gpurun_out/r1_bench_simt.log 2>&1
echo "exit $?" >> gpurun_out/r1_bench_simt.log
tail -5 gpurun_out/r1_pytest_simt.log gpurun_out/r1_pytest_linear.log gpurun_out/r1_pytest_tc.log gpurun_out/r1_smoke.log gpurun_out/r1_bench.log
