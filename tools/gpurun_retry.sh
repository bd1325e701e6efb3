#!/bin/bash
# tools/gpurun_retry.sh <timeout> <log> <command...>: retry while the pod answers busy (exit 3), up to 10 times
t=$1; log=$2; shift 2
for k in 1 2 3 4 5 6 7 8 9 10; do
  /usr/local/graft/bin/gpurun --timeout $t -- "$@" > $log 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" $log; then exit $rc; fi
  sleep 90
done
exit 3
