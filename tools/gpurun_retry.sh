#!/bin/bash
# dev helper: gpurun with retries while the pod answers "busy" (exit 3) — usage: tools/gpurun_retry.sh <timeout> '<command>' [--gpus N]
T=$1; shift; CMD=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" --timeout "$T" -- "$CMD"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
