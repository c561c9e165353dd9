# usage: bash scripts/gpu_ncu_kernel.sh <kernel-regex> <out-tag> [frames]
set +e
mkdir -p gpurun_out
timeout 300 python scripts/profile_step.py ${3:-1024} > gpurun_out/${2}_plain.log 2>&1 || { tail -5 gpurun_out/${2}_plain.log; exit 1; }
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$1" -o gpurun_out/${2} -f python scripts/profile_step.py ${3:-1024} > gpurun_out/${2}_ncu.log 2>&1
tail -3 gpurun_out/${2}_ncu.log
