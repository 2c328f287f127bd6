#!/bin/bash
# usage: tools/gpu.sh <timeout-seconds> '<command>'   -- gpurun with retries while the pod has no free slot (exit 3)
t=$1; shift
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  /usr/local/graft/bin/gpurun --timeout "$t" -- "$@"
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q '"status": "transient"' /root/repo/gpurun_out/.last_call.json 2>/dev/null; then exit $rc; fi
  sleep 60
done
exit 3
