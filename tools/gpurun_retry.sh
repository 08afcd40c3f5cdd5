#!/bin/bash
# tools/gpurun_retry.sh [gpurun options] -- '<command>' : retry while the pod answers "no box free" (exit 3), every 2 minutes, up to 40 times
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
