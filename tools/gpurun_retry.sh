#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> <timeout-seconds> '<command>'  -- retries while the pod answers busy (rc 2 "another call" / rc 3)
log=$1; to=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$to" -- "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "gpurun rc=$rc" >> "$log"; exit $rc; fi
  sleep 60
done
echo "gpurun gave up (busy)" >> "$log"; exit 3
