#!/bin/bash
# twelfth GPU call: per-tile / per-iteration / per-launch cost of the iteration kernel; tiled warp kernel timing + bit-equality
mkdir -p gpurun_out
timeout 300 python tools/gpu_probe_tile_cost.py > gpurun_out/r2l_tile_cost.log 2>&1; cat gpurun_out/r2l_tile_cost.log
timeout 300 python -m pytest tests/test_tvl1_gpu.py -x -q -m gpu -k "tiled_warp or cuda_semantics" > gpurun_out/r2l_pytest.log 2>&1; tail -5 gpurun_out/r2l_pytest.log
timeout 300 python tools/gpu_probe_r2.py 0:8:0 0:8:3 > gpurun_out/r2l_probe.log 2>&1; cat gpurun_out/r2l_probe.log
