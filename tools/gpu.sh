#!/bin/bash
# usage: tools/gpu.sh <timeout-seconds> '<command>'   - gpurun with retries while the pod answers "transient" / busy
T=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@" > /tmp/gpurun_last.txt 2>&1
  rc=$?
  if grep -q "status=transient\|rc=3\|no box\|busy" /tmp/gpurun_last.txt && ! grep -q "status=ok" /tmp/gpurun_last.txt; then
    sleep 60
    continue
  fi
  break
done
cat /tmp/gpurun_last.txt
exit $rc
