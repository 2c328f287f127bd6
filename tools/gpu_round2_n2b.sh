#!/bin/bash
# two-GPU call: one-channel native NCCL gather vs torch gather (TV-L1 and Farneback)
mkdir -p gpurun_out
for w in tvl1 farneback; do for g in native torch; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --workload $w --gather $g --no-cpu --no-extras > gpurun_out/r2n2b_${w}_$g.json 2> gpurun_out/r2n2b_${w}_$g.err
  echo "workload=$w gather=$g rc=$?"; python - <<PY
import json
for l in open("gpurun_out/r2n2b_${w}_$g.json"):
    if l.startswith("{"):
        d=json.loads(l); print(d["value"], d["ms_per_step"], d["config"].get("gather"))
PY
done; done
