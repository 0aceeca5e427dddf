#!/bin/bash
# scratch/gpu_retry.sh TIMEOUT 'command' : gpurun with retries while the pod is busy (exit code 3 / transient)
T=$1; shift
for i in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient\|status=busy\|rc=None"; then sleep 90; continue; fi
  echo "$out"; exit 0
done
echo "gave up"; exit 3
