#!/bin/bash
# gpurun with retries while the pod is busy (exit code 3 = nothing charged).
# usage: [GPUS=2] tools/gpu_retry.sh <timeout_s> '<command>'
T=$1; shift
G=${GPUS:-1}
for i in $(seq 1 40); do
  if [ "$G" -gt 1 ]; then
    /usr/local/graft/bin/gpurun --gpus "$G" --timeout "$T" -- "$@"
  else
    /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"
  fi
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 75
done
exit 3
