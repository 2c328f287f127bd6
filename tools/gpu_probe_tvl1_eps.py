"""TV-L1 at the reference's create() defaults (eps = 0.01): time, samples, per-class breakdown."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import opencv_contrib_b200 as ocb
from oracle import synth
dev = torch.device("cuda:0")
for kind in ("smooth", "const"):
    I0, I1, gt = synth.make_pair(1080, 1920, seed=0, kind=kind)
    d0, d1 = torch.from_numpy(I0).to(dev), torch.from_numpy(I1).to(dev)
    alg = ocb.OpticalFlowDual_TVL1_create()
    flow = torch.empty((1080, 1920, 2), dtype=torch.float32, device=dev)
    for _ in range(2): alg.calc(d0, d1, flow)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5): alg.calc(d0, d1, flow)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 5 * 1e3
    alg.resetStats(); alg.calc(d0, d1, flow); torch.cuda.synchronize()
    st = alg.getStats()
    print(kind, "%.2f ms/pair" % ms, "iterations", st["iterations_run"], {k: v["launches"] for k, v in st["classes"].items()}, flush=True)
    alg.setProfiling(True); alg.resetStats(); alg.calc(d0, d1, flow); torch.cuda.synchronize()
    print("   profiled ms:", {k: round(v["ms"], 3) for k, v in alg.getStats()["classes"].items()}, flush=True)
