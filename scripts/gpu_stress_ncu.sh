# launch list + full-set capture of the aggregation kernels on the large-frame configuration (10 views x 16 persons, 64 frames)
set +e
mkdir -p gpurun_out
timeout 300 python scripts/agg_probe.py 64 ring10 16 > gpurun_out/sn_probe.log 2>&1; tail -5 gpurun_out/sn_probe.log
timeout 300 python scripts/profile_step.py 64 ring10 16 > gpurun_out/sn_plain.log 2>&1 || { tail -5 gpurun_out/sn_plain.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_stress.csv python scripts/profile_step.py 64 ring10 16 > gpurun_out/sn_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"gat_aggregate_large|cluster_kernel" -c 6 -o gpurun_out/prof_stress -f python scripts/profile_step.py 64 ring10 16 > gpurun_out/sn_ncu2.log 2>&1
tail -n 3 gpurun_out/sn_plain.log gpurun_out/sn_ncu1.log gpurun_out/sn_ncu2.log
