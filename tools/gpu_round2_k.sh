#!/bin/bash
# eleventh GPU call: MUFU throughput microbenchmark again (its first log was lost with the container); fused-iteration sweep of the
# default TV-L1 kernel now that halo-4 / halo-8 passes fill registers with LDS.128
mkdir -p gpurun_out
timeout 60 tools/ubench/ubench_mufu > gpurun_out/r2k_ubench_mufu.log 2>&1; cat gpurun_out/r2k_ubench_mufu.log
timeout 500 python tools/gpu_probe_r2.py 0:8 0:4 0:5 0:6 0:10 0:12 > gpurun_out/r2k_probe.log 2>&1; cat gpurun_out/r2k_probe.log
