#!/bin/bash
# seventeenth GPU call: whole GPU suite + default bench line with the tiled warp kernel and the multi-warp TMA issue as defaults
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2q_pytest.log 2>&1; tail -5 gpurun_out/r2q_pytest.log
timeout 900 python bench.py > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; tail -3 gpurun_out/r2q_bench.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2q_bench.json").read().strip().splitlines()[-1])
print("tvl1", d["value"], "e2e", d["e2e"]["value"], "1stream", d["roofline"].get("value_1stream"), d["roofline"]["all_classes_ms_per_pair"])
print("farneback", d["farneback"]["value"], d["farneback"]["e2e"]["value"]); print("4k", d["tvl1_4k"]["value"]); print("extras", d["extras"]); print("clocks", d["clocks"])
PY
