#!/bin/bash
# measured parity numbers of the round-2 full-size / golden tests
mkdir -p gpurun_out
timeout 900 python tools/gpu_parity_r2.py 2>&1 | tee gpurun_out/r2ab_parity.log
