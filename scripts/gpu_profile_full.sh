set +e
mkdir -p gpurun_out
timeout 300 python scripts/profile_step.py > gpurun_out/p4_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_v5.csv python scripts/profile_step.py > gpurun_out/p4_ncu1.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"gemm_split_tc2_kernel|gat_aggregate|gemm_small_m|lift_person|cluster_kernel|head_features|build_graph" -o gpurun_out/prof_r1_v5 -f python scripts/profile_step.py 1024 > gpurun_out/p4_ncu2.log 2>&1
tail -n 3 gpurun_out/p4_plain.log gpurun_out/p4_ncu1.log gpurun_out/p4_ncu2.log
