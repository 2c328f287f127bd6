#!/bin/bash
# nineteenth GPU call: descriptor prefetch; halo rounded to 4 for the last pass of a warp; engine streams per GPU 4 / 6 / 8
mkdir -p gpurun_out
timeout 300 python tools/gpu_probe_r2.py 0:8:0 > gpurun_out/r2s_probe.log 2>&1; cat gpurun_out/r2s_probe.log
B2F_DBG_TVL1_HALO4=1 timeout 300 python tools/gpu_probe_r2.py 0:8:0 > gpurun_out/r2s_probe_halo4.log 2>&1; cat gpurun_out/r2s_probe_halo4.log
for n in 4 6 8; do
  timeout 300 python bench.py --workload tvl1 --no-cpu --no-extras --streams $n --steps 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('streams', d['config']['streams_per_gpu'], 'value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2))"
done 2>&1 | tee gpurun_out/r2s_streams.log
