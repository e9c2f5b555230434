#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> [--gpus N] <command...>   (retries while the pod answers "busy", exit code 3)
T=$1; shift
OPTS=()
if [ "$1" = "--gpus" ]; then OPTS=(--gpus "$2"); shift 2; fi
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" "${OPTS[@]}" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
