#!/bin/bash
# ninth GPU call: MUFU throughput microbenchmark; TMA-store epilogue (kernel_path 10 / 11) sanity + timing
mkdir -p gpurun_out
timeout 60 tools/ubench/ubench_mufu > gpurun_out/r2i_ubench_mufu.log 2>&1; cat gpurun_out/r2i_ubench_mufu.log
timeout 200 python tools/gpu_cluster_sanity.py 0 10 11 > gpurun_out/r2i_sanity.log 2>&1; rc=$?; grep -E "path 1[01]|SANITY" gpurun_out/r2i_sanity.log | tail -9; echo "sanity rc=$rc"
if [ $rc -ne 0 ]; then tail -5 gpurun_out/r2i_sanity.log; exit 1; fi
timeout 400 python tools/gpu_probe_r2.py 0:8 10:8 11:8 > gpurun_out/r2i_probe.log 2>&1; cat gpurun_out/r2i_probe.log
