#!/bin/bash
# gpu_retry.sh <log> <timeout_s> <command...>: gpurun with retries while the pod answers "busy" (exit code 3: nothing charged).
log=$1; shift; to=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to "$@" > $log 2>&1
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 90
done
exit 3
