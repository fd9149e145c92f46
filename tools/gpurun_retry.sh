#!/bin/bash
# tools/gpurun_retry.sh <log> <gpurun args...> — retries a gpurun call while the pod answers "transient" (nothing charged).
log=$1; shift
for attempt in 1 2 3 4 5 6 7 8 9 10 11 12; do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  if ! grep -q "status=transient" "$log"; then break; fi
  sleep 90
done
