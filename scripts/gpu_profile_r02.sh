# One gpurun call: ncu launch list + --set full capture of one 1024-frame step (scripts/profile_step.py), round 2 kernel set.
# The raw-page CSV is exported on the box and the report itself dropped (gpurun merges at most 64 MiB back).
set +e
mkdir -p gpurun_out
timeout 300 python scripts/profile_step.py > gpurun_out/r02_plain.log 2>&1 || { tail -5 gpurun_out/r02_plain.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_1024frames.csv python scripts/profile_step.py > gpurun_out/r02_ncu1.log 2>&1
timeout 1500 ncu --set full --clock-control none --profile-from-start off -o /tmp/r02_full -f python scripts/profile_step.py 1024 > gpurun_out/r02_ncu2.log 2>&1
ncu -i /tmp/r02_full.ncu-rep --page raw --csv > gpurun_out/r02_full_raw.csv 2>/dev/null
ls -la /tmp/r02_full.ncu-rep gpurun_out/r02_full_raw.csv
tail -n 2 gpurun_out/r02_plain.log gpurun_out/r02_ncu1.log gpurun_out/r02_ncu2.log
