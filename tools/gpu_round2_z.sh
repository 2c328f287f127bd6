#!/bin/bash
# final single-GPU validation of round 2: whole GPU suite, smoke(), default bench line (kept as profiles/r02_bench_default_n1.json)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2z_pytest.log 2>&1; tail -4 gpurun_out/r2z_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1; tail -4 gpurun_out/r2z_smoke.log
timeout 900 python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; tail -2 gpurun_out/r2z_bench.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2z_bench.json").read().strip().splitlines()[-1])
print("tvl1", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), "1stream", round(d["roofline"].get("value_1stream"), 2), d["roofline"]["all_classes_ms_per_pair"])
print("roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "dram_frac", "ncu_issue_active", "avg_launch_us", "share_of_step") if k in d["roofline"]})
print("farneback", round(d["farneback"]["value"], 1), round(d["farneback"]["e2e"]["value"], 1)); print("4k", round(d["tvl1_4k"]["value"], 2), round(d["tvl1_4k"]["e2e"]["value"], 2)); print("extras", d["extras"]); print("clocks", d["clocks"]); print("cpu", d["cpu_baseline"]["value"], d["farneback"]["cpu_baseline"]["value"])
PY
