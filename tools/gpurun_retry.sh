#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> <logname> <command...>   -- retries while the pod answers busy (exit 3), nothing is charged for those
TO=$1; LOG=$2; shift 2
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $TO -- "$@" > gpurun_out/${LOG}_call.log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
