set +e
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -k "degenerate or end_to_end" > gpurun_out/p_pytest.log 2>&1; echo "exit $?" >> gpurun_out/p_pytest.log
timeout 300 python scripts/profile_step.py > gpurun_out/p_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py > gpurun_out/p_ncu1.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"gemm_split_tc_kernel|gat_aggregate_kernel|encode_persons" -o gpurun_out/prof_r1 python scripts/profile_step.py 256 > gpurun_out/p_ncu2.log 2>&1
tail -3 gpurun_out/p_pytest.log gpurun_out/p_plain.log gpurun_out/p_ncu1.log gpurun_out/p_ncu2.log
ls -la gpurun_out
