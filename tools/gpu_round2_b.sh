#!/bin/bash
# second GPU call of round 2: cluster-kernel sanity, the whole GPU suite, smoke, the default bench line, path probe
mkdir -p gpurun_out
timeout 240 python tools/gpu_cluster_sanity.py > gpurun_out/r2_cluster_sanity.log 2>&1
rc=$?; tail -14 gpurun_out/r2_cluster_sanity.log; echo "cluster sanity rc=$rc"
if [ $rc -ne 0 ]; then export B2F_SKIP_CLUSTER=1; fi
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_all.log 2>&1
tail -15 gpurun_out/r2_pytest_all.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -3 gpurun_out/r2_smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
tail -c 6000 gpurun_out/r2_bench_default.json; tail -5 gpurun_out/r2_bench_default.err
if [ $rc -eq 0 ]; then timeout 600 python tools/gpu_probe_r2.py > gpurun_out/r2_probe_b.log 2>&1; cat gpurun_out/r2_probe_b.log; fi
