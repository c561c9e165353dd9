# single-frame step: launch list (cold-cache kernel times of one live frame)
set +e
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_1frame.csv python scripts/profile_step.py 1 > gpurun_out/l1_ncu.log 2>&1
tail -n 2 gpurun_out/l1_ncu.log
