#!/bin/bash
# usage: gpurun_retry.sh <logfile> <gpurun args...> : retries while the pod answers "transient" (no box / slot free; nothing charged)
LOG=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  if ! grep -q "status=transient" "$LOG"; then exit 0; fi
  sleep 90
done
