#!/bin/bash
# twentieth GPU call: A/B on one box: descriptor prefetch on / off, halo rounded to 4 on / off (two rounds each)
mkdir -p gpurun_out
for r in 1 2; do
  timeout 300 python tools/gpu_probe_r2.py 0:8:0 2>&1 | sed 's/^/base      /'
  B2F_DBG_TVL1_PREFETCH=1 timeout 300 python tools/gpu_probe_r2.py 0:8:0 2>&1 | sed 's/^/prefetch  /'
  B2F_DBG_TVL1_HALO4=1 timeout 300 python tools/gpu_probe_r2.py 0:8:0 2>&1 | sed 's/^/halo4     /'
  B2F_DBG_TVL1_HALO4=1 B2F_DBG_TVL1_PREFETCH=1 timeout 300 python tools/gpu_probe_r2.py 0:8:0 2>&1 | sed 's/^/both      /'
done | tee gpurun_out/r2t_ab.log
