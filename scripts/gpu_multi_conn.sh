# usage: bash scripts/gpu_multi_conn.sh <n_gpus> <tag> - 3 and 2 compute lanes with CUDA_DEVICE_MAX_CONNECTIONS=32 (default 8: more streams than
# hardware queues alias onto the same queue and pick up false dependencies)
set +e
N=${1:-4}; TAG=${2:-mc}
mkdir -p gpurun_out
for L in 3 2; do
  CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$L bench.py --gpus $N --steps 20 --warmup 3 --lanes $L --cpu-budget 1 --latency-frames 0 > gpurun_out/${TAG}_n${N}_l$L.log 2>&1
  grep -h '^{' gpurun_out/${TAG}_n${N}_l$L.log | python -c "
import json,sys
for line in sys.stdin:
    l=json.loads(line); print('conn32 lanes', l['lanes'], 'value', round(l['value']), 'ms', round(l['ms_per_step'],3), 'one_lane_ms', round(l['one_lane']['ms_per_step'],3), 'e2e', round(l['e2e']['value']), 'e2e_ms', round(l['e2e']['ms_per_step'],3), l['clocks']['reasons'])"
done
