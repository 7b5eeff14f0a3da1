#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> <command...>   — retries while the pod answers "busy" (exit 3), nothing is charged for those
T=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@"; rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[retry $i] busy, sleeping 120 s"; sleep 120
done
exit 3
