#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit 3: nothing charged).  usage: [GPUS=N] tools/gpurun_retry.sh <timeout_s> '<command>'
T=$1; shift
G=""; [ -n "$GPUS" ] && G="--gpus $GPUS"
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun $G --timeout $T -- "$@"
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 90
done
exit 3
