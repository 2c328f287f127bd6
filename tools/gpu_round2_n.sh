#!/bin/bash
# fourteenth GPU call: per-tile phase clocks of the default iteration kernel
mkdir -p gpurun_out
timeout 300 python tools/gpu_probe_tile_clocks.py > gpurun_out/r2n_clocks.log 2>&1; cat gpurun_out/r2n_clocks.log
