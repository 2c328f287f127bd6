#!/bin/bash
# thirteenth GPU call: two-phase tiled warp kernel (slow pixels queued): bit-equality, timing per level size, 1080p pair rate
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_tvl1_gpu.py -x -q -m gpu -k "tiled_warp" > gpurun_out/r2m_pytest.log 2>&1; tail -5 gpurun_out/r2m_pytest.log
timeout 300 python tools/gpu_probe_tile_cost.py warp > gpurun_out/r2m_tile_cost.log 2>&1; cat gpurun_out/r2m_tile_cost.log
timeout 300 python tools/gpu_probe_r2.py 0:8:0 0:8:3 > gpurun_out/r2m_probe.log 2>&1; cat gpurun_out/r2m_probe.log
