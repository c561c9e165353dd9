set +e
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "triangulation_full" --timeout 300 > gpurun_out/x_pytest.log 2>&1; echo "exit $?" >> gpurun_out/x_pytest.log
timeout 300 python bench.py --workload triangulation --steps 10 > gpurun_out/x_tri.log 2>&1; echo "exit $?" >> gpurun_out/x_tri.log
timeout 600 python bench.py --config arp3 --persons 8 --steps 10 --cpu-budget 3 --latency-frames 0 > gpurun_out/x_arp3.log 2>&1; echo "exit $?" >> gpurun_out/x_arp3.log
tail -5 gpurun_out/x_pytest.log; tail -2 gpurun_out/x_tri.log | cut -c1-900; tail -2 gpurun_out/x_arp3.log | cut -c1-1200
