#!/bin/bash
# usage: scripts/gpurun_retry.sh TIMEOUT 'command' -- retries while the pod answers "busy / draining" (exit 3)
T=$1; shift
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
