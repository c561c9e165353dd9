set +e
mkdir -p gpurun_out
timeout 300 python scripts/agg_probe.py 1024 panoptic 4 0 > gpurun_out/ab_frame.log 2>&1
timeout 300 python scripts/agg_probe.py 1024 panoptic 4 2 > gpurun_out/ab_large.log 2>&1
tail -n 7 gpurun_out/ab_frame.log gpurun_out/ab_large.log
