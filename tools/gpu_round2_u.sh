#!/bin/bash
# A/B on one box: tiled warp kernel at 4 blocks / SM (64 registers) vs 3 blocks / SM (77 registers)
mkdir -p gpurun_out
for r in 1 2; do timeout 300 python tools/gpu_probe_r2.py 0:8:0 0:8:4; done 2>&1 | tee gpurun_out/r2u_ab.log
