"""One Brox solve without CUDA graphs (for an ncu launch list): python tools/prof_brox.py H W"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import opencv_contrib_b200 as ocb
from oracle import synth
h, w = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda:0")
I0, I1, _ = synth.make_pair(h, w, seed=0, kind="smooth", dtype="f32")
d0, d1 = torch.from_numpy(I0).to(dev), torch.from_numpy(I1).to(dev)
alg = ocb.BroxOpticalFlow_create(0.197, 50.0, 0.8, 10, 77, 10)
alg.setEngineOption("use_graph", 0)
flow = torch.empty((h, w, 2), dtype=torch.float32, device=dev)
alg.calc(d0, d1, flow)
torch.cuda.synchronize()
print("done", alg.getStats()["launches"])
