#!/bin/bash
# tools/gpurun_retry.sh <timeout-seconds> <command...>: gpurun with retries while the pod answers "busy" (nothing is charged then)
T=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient\|rc=3\|retry in a few minutes"; then sleep 90; continue; fi
  echo "$out"; exit 0
done
echo "$out"; echo "[gpurun_retry] gave up"; exit 3
