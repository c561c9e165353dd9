# One gpurun call: ncu launch lists + --set full captures with the round's final kernels - one 1024-frame inference step
# (scripts/profile_step.py) and one 15-graph optimisation step (scripts/profile_train_step.py). Each program runs once without
# ncu first. The raw pages are exported on the box and the reports dropped (gpurun merges at most 64 MiB back).
set +e
mkdir -p gpurun_out
timeout 300 python scripts/profile_step.py > gpurun_out/r02f_plain.log 2>&1 || { tail -5 gpurun_out/r02f_plain.log; exit 1; }
timeout 300 python scripts/profile_train_step.py > gpurun_out/r02f_train_plain.log 2>&1 || { tail -5 gpurun_out/r02f_train_plain.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02f_launches_1024frames.csv python scripts/profile_step.py > gpurun_out/r02f_ncu1.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02f_launches_train_step.csv python scripts/profile_train_step.py > gpurun_out/r02f_ncu1t.log 2>&1
timeout 1500 ncu --set full --clock-control none --profile-from-start off -o /tmp/r02f_full -f python scripts/profile_step.py 1024 > gpurun_out/r02f_ncu2.log 2>&1
ncu -i /tmp/r02f_full.ncu-rep --page raw --csv > gpurun_out/r02f_full_raw.csv 2>/dev/null
timeout 1500 ncu --set full --clock-control none --profile-from-start off -o /tmp/r02f_train_full -f python scripts/profile_train_step.py > gpurun_out/r02f_ncu2t.log 2>&1
ncu -i /tmp/r02f_train_full.ncu-rep --page raw --csv > gpurun_out/r02f_train_full_raw.csv 2>/dev/null
ls -la /tmp/r02f_full.ncu-rep /tmp/r02f_train_full.ncu-rep gpurun_out/r02f_full_raw.csv gpurun_out/r02f_train_full_raw.csv
tail -n 2 gpurun_out/r02f_plain.log gpurun_out/r02f_train_plain.log gpurun_out/r02f_ncu1.log gpurun_out/r02f_ncu2.log gpurun_out/r02f_ncu2t.log
