"""Quick check of the thread-block-cluster TV-L1 kernels (kernel_path 6 / 7) against the default path on a small
problem; run under `timeout` before the full suite so that a protocol bug (hang) costs seconds, not the budget."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import opencv_contrib_b200 as ocb
from oracle import synth
dev = torch.device("cuda:0")
PATHS = tuple(int(a) for a in sys.argv[1:]) or (0, 6, 7, 8, 9)  # path 0 first: it is the comparison base
ok = True
for (h, w) in [(203, 277), (540, 960)]:
    I0, I1, _ = synth.make_pair(h, w, seed=3, kind="smooth")
    d0, d1 = torch.from_numpy(I0).to(dev), torch.from_numpy(I1).to(dev)
    outs = {}
    for path in PATHS:
        for K in (8, 3):
            alg = ocb.OpticalFlowDual_TVL1_create(nscales=3, warps=2, epsilon=0.0, iterations=23)
            alg.setEngineOption("kernel_path", path)
            alg.setEngineOption("fused_iters", K)
            alg.setEngineOption("use_graph", 0)
            outs[(path, K)] = alg.calc(d0, d1).cpu().numpy()
            torch.cuda.synchronize()
            same = np.array_equal(outs[(path, K)], outs[(0, 8)])
            print(h, w, "path", path, "K", K, "bit-equal to path 0:", same, "max diff", float(np.abs(outs[(path, K)] - outs[(0, 8)]).max()), flush=True)
            ok = ok and same
print("CLUSTER_SANITY", "OK" if ok else "MISMATCH")
sys.exit(0 if ok else 1)
