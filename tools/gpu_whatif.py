import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import opencv_contrib_b200 as ocb
from oracle import synth
dev = torch.device("cuda:0")
I0, I1, gt = synth.make_pair(1080, 1920, seed=0, kind="smooth")
d0, d1 = torch.from_numpy(I0).to(dev), torch.from_numpy(I1).to(dev)
names = {0: "normal", 11: "no __syncthreads in loop", 12: "no SFU in dual", 13: "no shuffles"}
for K in (6, 8):
    for path in (0, 11, 12, 13):
        alg = ocb.OpticalFlowDual_TVL1_create(nscales=5, warps=10, epsilon=0.0, iterations=30)
        alg.setEngineOption("kernel_path", path); alg.setEngineOption("fused_iters", K)
        flow = torch.empty((1080, 1920, 2), dtype=torch.float32, device=dev)
        for _ in range(2): alg.calc(d0, d1, flow)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): alg.calc(d0, d1, flow)
        e1.record(); torch.cuda.synchronize()
        print("K=%d %-28s %.2f ms/pair" % (K, names[path], e0.elapsed_time(e1) / 5), flush=True)
