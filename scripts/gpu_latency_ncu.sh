# single-frame step: launch list, full-set capture of the weight-streaming MLP kernel, aggregation A/B (frame-resident vs large-frame kernel)
set +e
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_1frame.csv python scripts/profile_step.py 1 > gpurun_out/l1_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"gemm_small_m" -c 3 -o gpurun_out/prof_small_m -f python scripts/profile_step.py 1 > gpurun_out/l1_ncu2.log 2>&1
timeout 300 python scripts/agg_probe.py 1 panoptic 4 0 > gpurun_out/ab1_frame.log 2>&1
timeout 300 python scripts/agg_probe.py 1 panoptic 4 2 > gpurun_out/ab1_large.log 2>&1
tail -n 2 gpurun_out/l1_ncu.log gpurun_out/l1_ncu2.log; tail -n 6 gpurun_out/ab1_frame.log gpurun_out/ab1_large.log
