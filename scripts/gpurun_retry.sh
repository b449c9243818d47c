#!/bin/bash
# usage: [GPURUN_FLAGS="--gpus 2"] scripts/gpurun_retry.sh <timeout_s> '<command>'
# retries while the pod answers "busy" (exit 3) -- nothing is charged then
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun $GPURUN_FLAGS --timeout "$T" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
