#!/bin/bash
# usage: scripts/gpu_retry_n.sh <gpus> <timeout_s> '<command>'
G=$1; shift; T=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --gpus "$G" --timeout "$T" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
