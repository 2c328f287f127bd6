import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import opencv_contrib_b200 as ocb
from oracle import synth
path = int(sys.argv[1]); K = int(sys.argv[2]) if len(sys.argv) > 2 else 6
dev = torch.device("cuda:0")
I0, I1, gt = synth.make_pair(203, 277, seed=1, kind="smooth")
kw = dict(nscales=2, warps=1, epsilon=0.0, iterations=6)
outs = []
for p in (1, path):
    alg = ocb.OpticalFlowDual_TVL1_create(**kw)
    alg.setEngineOption("kernel_path", p); alg.setEngineOption("fused_iters", K); alg.setEngineOption("use_graph", 0)
    f = alg.calc(torch.from_numpy(I0).to(dev), torch.from_numpy(I1).to(dev)); torch.cuda.synchronize()
    outs.append(f.cpu().numpy())
print("path", path, "K", K, "bit-equal:", np.array_equal(outs[0], outs[1]), float(np.abs(outs[0]-outs[1]).max()))
