# usage: bash scripts/gpu_multi_lanes.sh <n_gpus> <tag>  - the headline bench on n GPUs with 1, 2 and 3 compute lanes (short form: no CPU arm, no latency loop)
set +e
N=${1:-4}; TAG=${2:-ml}
mkdir -p gpurun_out
nproc > gpurun_out/${TAG}_nproc.txt
for L in 1 3 2; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$L bench.py --gpus $N --steps 20 --warmup 3 --lanes $L --cpu-budget 1 --latency-frames 0 > gpurun_out/${TAG}_n${N}_l$L.log 2>&1
  grep -h '^{' gpurun_out/${TAG}_n${N}_l$L.log | python -c "
import json,sys
for line in sys.stdin:
    l=json.loads(line); print('lanes', l['lanes'], 'value', round(l['value']), 'ms', round(l['ms_per_step'],3), 'one_lane_ms', round(l['one_lane']['ms_per_step'],3), 'e2e', round(l['e2e']['value']), 'e2e_ms', round(l['e2e']['ms_per_step'],3), l['clocks']['reasons'])"
done
cat gpurun_out/${TAG}_nproc.txt
