#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit 3); usage: gpurun_retry.sh <timeout> <tag> <command...>
TO=$1; TAG=$2; shift 2
for i in 1 2 3 4 5 6 7 8; do
  /usr/local/graft/bin/gpurun --timeout $TO -- "$@" > gpurun_out/${TAG}_call.log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
