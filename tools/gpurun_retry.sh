#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit 3 / transient: nothing charged).
# usage: tools/gpurun_retry.sh [gpurun options] -- <command>
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > /tmp/gpurun_last.log 2>&1
  rc=$?
  if grep -q "status=transient" /tmp/gpurun_last.log || [ $rc -eq 3 ]; then sleep 120; continue; fi
  cat /tmp/gpurun_last.log; exit $rc
done
cat /tmp/gpurun_last.log; exit 3
