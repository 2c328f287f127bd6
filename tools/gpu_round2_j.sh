#!/bin/bash
# tenth GPU call: Farneback lane-transposed gather (aux_path 5): bit-equality on odd sizes, timing, one ncu capture
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_farneback_gpu.py -x -q -m gpu -k "variants" > gpurun_out/r2j_pytest.log 2>&1; tail -3 gpurun_out/r2j_pytest.log
timeout 300 python tools/gpu_probe_farn_r2.py 0 5 > gpurun_out/r2j_probe.log 2>&1; cat gpurun_out/r2j_probe.log
