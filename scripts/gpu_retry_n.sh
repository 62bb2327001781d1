#!/bin/bash
# usage: scripts/gpu_retry_n.sh <gpus> <timeout-seconds> '<command>'   -- multi-GPU variant of gpu_retry.sh (retries on "busy")
N=$1; T=$2; shift 2
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --gpus "$N" --timeout "$T" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 60
done
exit 3
