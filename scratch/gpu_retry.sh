#!/bin/bash
# usage: gpu_retry.sh <timeout> <command...>  -- retries gpurun while the pod answers "busy" (nothing charged)
t=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $t "$@" > /tmp/gpurun_last.log 2>&1
  rc=$?
  if grep -q "status=transient" /tmp/gpurun_last.log || [ $rc -eq 3 ]; then sleep 90; continue; fi
  break
done
tail -15 /tmp/gpurun_last.log
exit $rc
