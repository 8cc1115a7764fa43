#!/bin/bash
# usage: scripts/gpurun_retry.sh <logfile> <gpurun args...>   -- retries while the pod answers "busy" (exit code 3: nothing charged)
log=$1; shift
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "[retry] attempt $attempt rc=$rc" >> "$log"; exit $rc; fi
  sleep 60
done
echo "[retry] gave up" >> "$log"; exit 3
