#!/bin/bash
# fifteenth GPU call: row-skewed iteration kernel (kernel_path 12): bit-equality, timing
mkdir -p gpurun_out
B2F_SKIP_CLUSTER=1 timeout 600 python -m pytest tests/test_tvl1_gpu.py -x -q -m gpu -k "bit_identical_to_unfused" > gpurun_out/r2o_pytest.log 2>&1; tail -5 gpurun_out/r2o_pytest.log
timeout 300 python tools/gpu_probe_r2.py 0:8:3 12:8:3 12:12:3 > gpurun_out/r2o_probe.log 2>&1; cat gpurun_out/r2o_probe.log
