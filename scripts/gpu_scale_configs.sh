# usage: bash scripts/gpu_scale_configs.sh <n_gpus> [arp3|ring10|panoptic ...]   (the command of one `gpurun --gpus N` call)
# BASELINE.json configs[2] (ARP 3-view x 8 persons, 1 M frames over the GPUs, chunks of 4096 frames) and configs[4]
# (10-view x 16-person stress frames, 64 per GPU and step), one JSON line each under gpurun_out/.
set +e
N=${1:-1}; shift
mkdir -p gpurun_out
run() {  # tag, bench args...
  local tag=$1; shift
  if [ "$N" = "1" ]; then
    timeout 900 python bench.py --gpus 1 "$@" > gpurun_out/r2_scale_${tag}_n1.json 2> gpurun_out/r2_scale_${tag}_n1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N "$@" > gpurun_out/r2_scale_${tag}_n$N.json 2> gpurun_out/r2_scale_${tag}_n$N.err
  fi
  echo "$tag n=$N exit $?"; tail -c 600 gpurun_out/r2_scale_${tag}_n$N.json | head -c 300; echo
}
for w in "${@:-arp3 ring10}"; do
  case $w in
    arp3) run arp3 --config arp3 --persons 8 --frames 4096 --total-frames 1000000 --warmup 3 --latency-frames 0 --cpu-budget 5 ;;
    ring10) run ring10 --config ring10 --persons 16 --frames 64 --steps 10 --warmup 3 --latency-frames 0 --cpu-budget 5 ;;
    panoptic) run panoptic --steps 20 --warmup 3 --latency-frames 0 --cpu-budget 5 ;;
  esac
done
