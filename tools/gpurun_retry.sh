#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3).  Usage: tools/gpurun_retry.sh <logfile> <gpurun args...>
LOG=$1; shift
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun "$@" > $LOG 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
