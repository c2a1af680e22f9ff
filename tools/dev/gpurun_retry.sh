#!/bin/bash
# gpurun_retry.sh TIMEOUT CMD...: gpurun with retries while the pod answers "no box / slot free" (exit 3). Dev tool.
t=$1; shift
for attempt in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout "$t" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
