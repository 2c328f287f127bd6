#!/bin/bash
# usage: [GPUS=2] tools/gpu.sh <timeout-seconds> '<command>'   -- gpurun with retries while the pod has no free slot
t=$1; shift
g=${GPUS:-1}
for i in 1 2 3 4 5 6 7 8 9 10 11 12 13 14 15 16; do
  if [ "$g" = "1" ]; then /usr/local/graft/bin/gpurun --timeout "$t" -- "$@"; else /usr/local/graft/bin/gpurun --gpus "$g" --timeout "$t" -- "$@"; fi
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q '"status": "transient"' /root/repo/gpurun_out/.last_call.json 2>/dev/null; then exit $rc; fi
  sleep 45
done
exit 3
