#!/bin/bash
# usage: gp.sh TIMEOUT_S [--gpus N] -- 'command'   : gpurun with retries while the pod answers "busy" (exit code 3)
t=$1; shift
extra=()
while [ "$1" != "--" ]; do extra+=("$1"); shift; done
shift
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$t" "${extra[@]}" -- "$1"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[gp.sh] busy (attempt $attempt), retrying in 90 s"
  sleep 90
done
exit 3
