"""Small-size pass over every engine and adjacent component (run under compute-sanitizer)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, cv2
import opencv_contrib_b200 as ocb
from oracle import synth
dev = torch.device("cuda:0")
I0, I1, _ = synth.make_pair(141, 203, seed=3, kind="smooth")
d0, d1 = torch.from_numpy(I0).to(dev), torch.from_numpy(I1).to(dev)
f0, f1 = d0.float() / 255, d1.float() / 255
for name, alg, a, b in [
        ("tvl1", ocb.OpticalFlowDual_TVL1_create(nscales=3, warps=2, epsilon=0.0, iterations=9), d0, d1),
        ("tvl1-eps", ocb.OpticalFlowDual_TVL1_create(nscales=2, warps=2, iterations=12), d0, d1),
        ("farneback", ocb.FarnebackOpticalFlow_create(numLevels=3, numIters=2), d0, d1),
        ("brox", ocb.BroxOpticalFlow_create(0.197, 50.0, 0.8, 2, 10, 7), f0, f1),
        ("brox-pp", ocb.BroxOpticalFlow_create(0.197, 50.0, 0.8, 2, 10, 3), f0, f1),
        ("brox-coop", ocb.BroxOpticalFlow_create(0.197, 50.0, 0.8, 3, 10, 5), f0, f1),
        ("tvl1-skewed", ocb.OpticalFlowDual_TVL1_create(nscales=2, warps=2, epsilon=0.0, iterations=8), d0, d1),
        ("denselk", ocb.DensePyrLKOpticalFlow_create(maxLevel=2, iters=4), d0, d1),
        ("denselk-generic", ocb.DensePyrLKOpticalFlow_create(winSize=(9, 7), maxLevel=1, iters=3), d0, d1)]:
    if name == "brox-pp":
        alg.setEngineOption("kernel_path", 2)
    if name == "brox-coop":
        alg.setEngineOption("kernel_path", 3)
    if name == "tvl1-skewed":
        alg.setEngineOption("kernel_path", 12)
    if len(sys.argv) > 1 and not any(name.startswith(a) for a in sys.argv[1:]):
        continue
    alg.setEngineOption("use_graph", 0)
    fl = alg.calc(a, b, torch.zeros((141, 203, 2), device=dev))
    u, v = alg.calcUV(a, b, torch.zeros((141, 203), device=dev), torch.zeros((141, 203), device=dev))
    torch.cuda.synchronize()
    print(name, "ok", float(fl.abs().mean()), flush=True)
fl = fl.contiguous()
mid = ocb.interpolateFrames(f0, f1, fl[..., 0].contiguous(), fl[..., 1].contiguous(), -fl[..., 0].contiguous(),
                            -fl[..., 1].contiguous(), 0.3)
mid2 = ocb.interpolateFrames(f0, f1, fl[..., 0].contiguous(), fl[..., 1].contiguous(), -fl[..., 0].contiguous(),
                             -fl[..., 1].contiguous(), 0.3, corrected=True)
torch.cuda.synchronize()
print("interp ok", float(mid.mean()), float(mid2.mean()), flush=True)
pts = cv2.goodFeaturesToTrack(I0, 200, 0.01, 0.0).reshape(-1, 2).astype(np.float32)
sp = ocb.SparsePyrLKOpticalFlow_create(maxLevel=2, iters=8)
n, s, e = sp.calc(d0, d1, torch.from_numpy(pts).to(dev), wantErr=True)
torch.cuda.synchronize()
print("sparse ok", int(s.sum()), flush=True)
vf = ocb.VideoFlow(ocb.FarnebackOpticalFlow_create(numLevels=2, numIters=2), 141, 203, depth=2, warm_start=True)
out = list(vf.run([I0, I1, I0, I1, I0]))
vf.close()
print("video ok", len(out), flush=True)
