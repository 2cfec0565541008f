#!/bin/bash
# tools/gpurun_retry.sh <log> <gpurun args...>: runs gpurun, retrying every 2 minutes while the pod answers "busy" (exit 3)
log=$1; shift
for attempt in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
