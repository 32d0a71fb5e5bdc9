#!/bin/bash
# usage: scripts/gpu_retry.sh <timeout> <logfile> <command...>   -- retries while the pod has no free GPU slot (rc 3)
t=$1; log=$2; shift 2
for k in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $t -- "$@" > $log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
