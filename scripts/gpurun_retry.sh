#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout-seconds> '<command>'  - retries while the pod has no free slot (exit 3), up to ~40 min
T=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
