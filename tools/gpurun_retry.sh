#!/bin/bash
# tools/gpurun_retry.sh [gpurun args...] -- gpurun, repeated while the pod answers "busy" (exit 3)
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@"; rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 90
done
exit 3
