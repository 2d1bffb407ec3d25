#!/bin/bash
# usage: tools/gpu_retry.sh <timeout-seconds> '<command>'  -- retries gpurun while the pod answers "transient/busy"
T=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient\|rc=3\|no box or slot"; then sleep 120; continue; fi
  echo "$out"; exit 0
done
echo "gave up: $out"; exit 3
