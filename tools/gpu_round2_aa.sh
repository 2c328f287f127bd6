#!/bin/bash
# Farneback: prefetch of the R1 gather window at block start (aux_path 6: to L2, 7: to L1) against the default, one box
mkdir -p gpurun_out
for r in 1 2; do timeout 300 python tools/gpu_probe_farn_r2.py 0 6 7; done 2>&1 | tee gpurun_out/r2aa_farn_prefetch.log
