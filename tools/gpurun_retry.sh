#!/bin/bash
# gpurun_retry.sh LOG [gpurun args...] -- retry while the pod answers "transient" (exit 3), at most 12 times
LOG=$1; shift
for i in $(seq 1 12); do
  gpurun "$@" > "$LOG" 2>&1; rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
