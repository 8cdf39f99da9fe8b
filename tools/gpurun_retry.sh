#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> <gpurun args...> ; retries while the pod answers busy (rc 3 / transient)
log=$1; shift
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if grep -q "status=transient\|nothing was charged" "$log"; then sleep 150; continue; fi
  exit $rc
done
exit 3
