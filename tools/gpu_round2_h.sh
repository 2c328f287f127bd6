#!/bin/bash
# eighth GPU call: neighbour-mbarrier kernel (kernel_path 9) sanity + timing + one ncu capture
mkdir -p gpurun_out
timeout 200 python tools/gpu_cluster_sanity.py 0 9 > gpurun_out/r2h_sanity.log 2>&1; rc=$?; grep -E "path 9|SANITY" gpurun_out/r2h_sanity.log | tail -6; echo "sanity rc=$rc"
if [ $rc -ne 0 ]; then tail -5 gpurun_out/r2h_sanity.log; exit 1; fi
timeout 400 python tools/gpu_probe_r2.py 0:8 9:8 9:6 9:12 > gpurun_out/r2h_probe.log 2>&1; cat gpurun_out/r2h_probe.log
timeout 300 ncu --set full --clock-control none --import-source on --launch-count 1 -f --kernel-name regex:k_tvl1_blocked_tma --launch-skip 165 -o gpurun_out/prof_tvl1_nb_r2h python tools/prof_one.py tvl1 9 1 8 > /dev/null 2>&1
ls -la gpurun_out/prof_tvl1_nb_r2h.ncu-rep
