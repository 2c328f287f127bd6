#!/bin/bash
# Farneback: register-blocked polynomial expansion (default) vs the round-1 kernel (aux_path 6): tests, A/B on one box
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_farneback_gpu.py -x -q -m gpu > gpurun_out/r2ae_pytest.log 2>&1; tail -4 gpurun_out/r2ae_pytest.log
for r in 1 2; do timeout 300 python tools/gpu_probe_farn_r2.py 6 0; done 2>&1 | tee gpurun_out/r2ae_polyexp.log
