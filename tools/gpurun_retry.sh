#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> [--gpus N] '<command>'   — retries while the pod answers "busy" (exit code 3)
t=$1; shift
extra=()
if [ "$1" = "--gpus" ]; then extra=(--gpus "$2"); shift 2; fi
for i in $(seq 1 25); do
  /usr/local/graft/bin/gpurun --timeout "$t" "${extra[@]}" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
