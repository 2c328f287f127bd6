#!/bin/bash
# Farneback: four-column sparse vertical blur (default) vs round-1 secondary kernels (aux_path 6): tests, A/B on one box
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_farneback_gpu.py -x -q -m gpu > gpurun_out/r2af_pytest.log 2>&1; tail -3 gpurun_out/r2af_pytest.log
timeout 300 python tools/gpu_probe_farn_r2.py 6 0 6 0 2>&1 | tee gpurun_out/r2af_blurv.log
