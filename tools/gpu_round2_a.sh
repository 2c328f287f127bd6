#!/bin/bash
# first GPU call of round 2: microbenchmarks, parity tests with the new kernels, path/K sweep, Farneback graph check
mkdir -p gpurun_out
timeout 120 tools/ubench/ubench_fp32x2_dsmem > gpurun_out/ubench1.log 2>&1
timeout 1200 python -m pytest tests/test_tvl1_gpu.py tests/test_farneback_gpu.py -x -q -m gpu > gpurun_out/r2_pytest_tvl1.log 2>&1
tail -5 gpurun_out/r2_pytest_tvl1.log
timeout 600 python tools/gpu_probe_r2.py > gpurun_out/r2_probe_a.log 2>&1
timeout 300 python tools/gpu_probe_farneback.py 2>&1 | tail -12 > gpurun_out/r2_probe_farn.log
cat gpurun_out/ubench1.log gpurun_out/r2_probe_a.log gpurun_out/r2_probe_farn.log
