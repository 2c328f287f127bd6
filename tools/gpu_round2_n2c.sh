#!/bin/bash
# two-GPU call: the driver's own command line for N = 2 (default bench: TV-L1 headline + farneback + 4K sub-records, native gather)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2n2c_bench.json 2> gpurun_out/r2n2c_bench.err
echo "rc=$?"; tail -3 gpurun_out/r2n2c_bench.err; python - <<'PY'
import json
for l in open("gpurun_out/r2n2c_bench.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("N", d["n_gpus"], "tvl1", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1), "gather", d["config"].get("gather"), "nccl", d.get("nccl", {}))
        print("farneback", round(d["farneback"]["value"], 1), "4k", round(d["tvl1_4k"]["value"], 2))
PY
