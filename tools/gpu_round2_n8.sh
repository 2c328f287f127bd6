#!/bin/bash
# N-GPU call (N from $1, default 8): native NCCL gather vs torch gather on the TV-L1 headline workload
N=${1:-8}
mkdir -p gpurun_out
for g in native torch; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 --workload tvl1 --gather $g --no-cpu --no-extras > gpurun_out/r2n${N}_tvl1_$g.json 2> gpurun_out/r2n${N}_tvl1_$g.err
  echo "N=$N gather=$g rc=$?"; python - <<PY
import json
for l in open("gpurun_out/r2n${N}_tvl1_$g.json"):
    if l.startswith("{"):
        d=json.loads(l); print(d["value"], d["ms_per_step"], d.get("gather"), d.get("nccl",{}).get("world_size"))
PY
done
