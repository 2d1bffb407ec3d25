#!/bin/bash
# usage: tools/gpu_retry.sh [-g N] <timeout-seconds> '<command>'  -- retries gpurun while the pod answers "transient/busy"
G=""
if [ "$1" = "-g" ]; then G="--gpus $2"; shift 2; fi
T=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun $G --timeout "$T" -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient\|rc=3\|no box or slot\|status=busy"; then sleep 120; continue; fi
  echo "$out"; exit 0
done
echo "gave up: $out"; exit 3
