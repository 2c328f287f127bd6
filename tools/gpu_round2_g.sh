#!/bin/bash
# seventh GPU call: two-group kernel sanity + timing
mkdir -p gpurun_out
timeout 300 python tools/gpu_cluster_sanity.py > gpurun_out/r2g_sanity.log 2>&1; rc=$?; grep -E "path 8|SANITY" gpurun_out/r2g_sanity.log | tail -6; echo "sanity rc=$rc"
timeout 400 python tools/gpu_probe_r2.py 0:8 8:8 8:12 8:6 > gpurun_out/r2g_probe.log 2>&1; cat gpurun_out/r2g_probe.log
timeout 300 ncu --set full --clock-control none --import-source on --launch-count 1 -f --kernel-name regex:k_tvl1_blocked_tma --launch-skip 165 -o gpurun_out/prof_tvl1_2g_r2g python tools/prof_one.py tvl1 8 1 8 > /dev/null 2>&1
ls -la gpurun_out/prof_tvl1_2g_r2g.ncu-rep
