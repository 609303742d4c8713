#!/bin/bash
# usage: gpurun_retry.sh <log> <gpurun args...>  — retries while the pod has no free slot (rc 3 / transient)
log=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  if grep -q "status=transient\|no box or slot" "$log"; then sleep 90; continue; fi
  break
done
