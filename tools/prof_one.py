"""Run a few calc() calls of one algorithm at 1080p without CUDA graphs (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import opencv_contrib_b200 as ocb
from oracle import synth

algo = sys.argv[1] if len(sys.argv) > 1 else "tvl1"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 5
n = int(sys.argv[3]) if len(sys.argv) > 3 else 1
path = int(sys.argv[4]) if len(sys.argv) > 4 else 0
dev = torch.device("cuda:0")
I0, I1, gt = synth.make_pair(1080, 1920, seed=0, kind="smooth")
d0, d1 = torch.from_numpy(I0).to(dev), torch.from_numpy(I1).to(dev)
if algo == "tvl1":
    alg = ocb.OpticalFlowDual_TVL1_create(nscales=5, warps=10, epsilon=0.0, iterations=30)
    alg.setEngineOption("fused_iters", K)
    alg.setEngineOption("use_graph", 0)
    alg.setEngineOption("kernel_path", path)
elif algo == "denselk":
    alg = ocb.DensePyrLKOpticalFlow_create()
else:
    alg = ocb.FarnebackOpticalFlow_create()
    alg.setEngineOption("use_graph", 0)
flow = torch.empty((1080, 1920, 2), dtype=torch.float32, device=dev)
for _ in range(n):
    alg.calc(d0, d1, flow)
torch.cuda.synchronize()
print("done", alg.getStats()["launches"])
