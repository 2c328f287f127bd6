#!/bin/bash
# Farneback batched rate against the number of engine streams per GPU
mkdir -p gpurun_out
for n in 8 12 16; do
  timeout 300 python bench.py --workload farneback --no-cpu --no-extras --streams $n --steps 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('streams', d['config']['streams_per_gpu'], 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1))"
done 2>&1 | tee gpurun_out/r2y_farn_streams.log
