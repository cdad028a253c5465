#!/bin/bash
# usage: scripts/gpurun_retry.sh <log> <gpurun args...>   — retries while the pod answers "transient" (nothing is charged for those)
log=$1; shift
for i in $(seq 1 40); do
  gpurun "$@" > "$log" 2>&1
  if grep -q "status=transient" "$log" || grep -q "rc=3" "$log"; then sleep 45; continue; fi
  break
done
