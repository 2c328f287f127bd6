#!/bin/bash
# A/B on one box: warp tile kernel, pixel loop unrolled by 2 at 3 blocks / SM (aux 4) vs the default (aux 0)
mkdir -p gpurun_out
for r in 1 2; do timeout 300 python tools/gpu_probe_r2.py 0:8:0 0:8:4; done 2>&1 | tee gpurun_out/r2ac_ab.log
