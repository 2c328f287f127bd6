#!/bin/bash
# second GPU call of round 2: the whole GPU suite, smoke, the default bench line and a launch list
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_all.log 2>&1
tail -15 gpurun_out/r2_pytest_all.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -3 gpurun_out/r2_smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
tail -c 6000 gpurun_out/r2_bench_default.json; tail -5 gpurun_out/r2_bench_default.err
timeout 600 python tools/gpu_probe_r2.py > gpurun_out/r2_probe_b.log 2>&1; cat gpurun_out/r2_probe_b.log
