# usage: bash scripts/gpu_multi.sh <n_gpus> <tag>
set +e
N=${1:-2}; TAG=${2:-mg}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_gpus.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --cpu-budget 3 > gpurun_out/${TAG}_bench_n$N.log 2>&1
echo "exit $?" >> gpurun_out/${TAG}_bench_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 1 --warmup 0 > gpurun_out/${TAG}_ref_n$N.log 2>&1
echo "exit $?" >> gpurun_out/${TAG}_ref_n$N.log
grep -h '^{' gpurun_out/${TAG}_bench_n$N.log | cut -c1-400; tail -3 gpurun_out/${TAG}_bench_n$N.log | cut -c1-300; grep -h '^{' gpurun_out/${TAG}_ref_n$N.log | cut -c1-300; tail -2 gpurun_out/${TAG}_ref_n$N.log | cut -c1-200
