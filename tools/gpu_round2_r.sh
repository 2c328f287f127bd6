#!/bin/bash
# eighteenth GPU call: grad plane holds the thresholding constant (no per-tile reciprocal in the iteration kernels): suite + probe
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2r_pytest.log 2>&1; tail -3 gpurun_out/r2r_pytest.log
timeout 300 python tools/gpu_probe_r2.py 0:8:0 0:8:0 > gpurun_out/r2r_probe.log 2>&1; cat gpurun_out/r2r_probe.log
