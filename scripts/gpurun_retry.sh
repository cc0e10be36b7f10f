#!/bin/bash
# usage: scripts/gpurun_retry.sh <logfile> <timeout> <command...>  -- retries while the pod answers "busy" (exit 3)
log=$1; shift; to=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "gpurun rc=$rc" >> $log; exit $rc; fi
  sleep 90
done
