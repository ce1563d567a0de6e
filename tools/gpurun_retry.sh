#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> <log> <command...>  -- retries while the pod answers "transient"
t=$1; log=$2; shift 2
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $t -- "$@" > $log 2>&1
  if ! grep -q "status=transient" $log; then break; fi
  sleep 45
done
tail -3 $log
