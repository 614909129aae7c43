#!/bin/bash
# usage: scripts/gpurun_retry.sh <gpus> <timeout> '<command>'  -- retries while the pod has no free slot
for attempt in 1 2 3 4 5 6 7 8 9 10 11 12; do
  out=$(/usr/local/graft/bin/gpurun --gpus "$1" --timeout "$2" -- "$3" 2>&1)
  echo "$out" | tail -60
  if echo "$out" | grep -q "status=ok\|status=fail\|status=timeout"; then exit 0; fi
  echo "--- attempt $attempt did not run; sleeping 150 s"
  sleep 150
done
