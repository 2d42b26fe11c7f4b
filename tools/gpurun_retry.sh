#!/bin/bash
# gpurun with retries while the pod has no free GPU slot (exit code 3 = nothing charged).
# usage: tools/gpurun_retry.sh <timeout_s> '<command>'
t=$1; shift
for k in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$t" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
