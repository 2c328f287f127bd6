#!/bin/bash
# Brox cooperative level kernel: bit-equality against the launch-per-step schedule, 720p timing A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_brox_lk_gpu.py -x -q -m gpu -k brox > gpurun_out/r2x_pytest.log 2>&1; tail -15 gpurun_out/r2x_pytest.log
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/r2x_brox.log
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import opencv_contrib_b200 as ocb
from oracle import synth
dev = torch.device("cuda:0")
I0, I1, _ = synth.make_pair(720, 1280, seed=0, kind="smooth", dtype="f32")
d0, d1 = torch.from_numpy(I0).to(dev), torch.from_numpy(I1).to(dev)
ref = None
for rnd in range(2):
    for path in (3, 0):
        alg = ocb.BroxOpticalFlow_create(0.197, 50.0, 0.8, 10, 77, 10)
        alg.setEngineOption("kernel_path", path)
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            for _ in range(3): out = alg.calc(d0, d1, None, s)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(s):
            e0.record(s)
            for _ in range(10): out = alg.calc(d0, d1, None, s)
            e1.record(s)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        o = out.cpu().numpy()
        if ref is None: ref = o
        print("kernel_path %d: %.3f ms/pair (%.1f pairs/s), %d launches, bit-equal to first: %s"
              % (path, ms, 1000 / ms, alg.getStats()["launches"], np.array_equal(o, ref)), flush=True)
PY
