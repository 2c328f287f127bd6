"""Phase clocks of the default TV-L1 iteration kernel (B2F_DBG_TVL1_CLOCKS: thread 0 of every CTA stamps clock64 around
wait / fill / set-up / iterations / epilogue of every tile; the library prints the per-tile averages to stderr)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["B2F_DBG_TVL1_CLOCKS"] = "1"
import numpy as np, torch
import opencv_contrib_b200 as ocb
from oracle import synth
dev = torch.device("cuda:0")
for (h, w) in ((1080, 1920), (864, 1536), (442, 786), (384, 48 * 148)):
    I0, I1, _ = synth.make_pair(h, w, seed=0, kind="smooth")
    d0, d1 = torch.from_numpy(I0).to(dev), torch.from_numpy(I1).to(dev)
    alg = ocb.OpticalFlowDual_TVL1_create(nscales=1, warps=2, epsilon=0.0, iterations=8)
    alg.setEngineOption("use_graph", 0)
    flow = torch.empty((h, w, 2), dtype=torch.float32, device=dev)
    for it in (None, 4, 0):
        if it is None: os.environ.pop("B2F_DBG_TVL1_ITERS", None)
        else: os.environ["B2F_DBG_TVL1_ITERS"] = str(it)
        alg.calc(d0, d1, flow); torch.cuda.synchronize()
    os.environ.pop("B2F_DBG_TVL1_ITERS", None)
