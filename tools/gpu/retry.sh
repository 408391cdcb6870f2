#!/bin/bash
# usage: tools/gpu/retry.sh <timeout> <logfile> <command...>   -- retries while gpurun answers "no slot" (exit 3)
to=$1; log=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$to" -- "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
