#!/bin/bash
# usage: tools/gpurun_retry.sh <log> <timeout> <command...>   — retries while the pod answers "transient" (nothing charged)
LOG=$1; TMO=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $TMO -- "$@" > $LOG 2>&1
  if ! grep -q "status=transient" $LOG; then break; fi
  sleep 120
done
tail -60 $LOG
