#!/bin/bash
# Retry gpurun until the pod has a slot: tools/gpu_retry.sh <log> <timeout> [--gpus N] -- '<command>'
log=$1; to=$2; shift; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to "$@" > $log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
