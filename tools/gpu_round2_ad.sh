#!/bin/bash
# compute-sanitizer over the small all-engine pass (memcheck) and over the TV-L1 kernels alone (racecheck: shared-memory hazards of
# the tiled warp kernel's queue, the multi-warp TMA issue and the split-phase variant)
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/memcheck_small.py > gpurun_out/r2ad_memcheck.log 2>&1; grep -E "ERROR SUMMARY|ok|Invalid|Error" gpurun_out/r2ad_memcheck.log | head -30
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/memcheck_small.py tvl1 > gpurun_out/r2ad_racecheck.log 2>&1; grep -E "RACECHECK SUMMARY|ok|hazard|Error" gpurun_out/r2ad_racecheck.log | head -30
