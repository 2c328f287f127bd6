"""First-contact probe (GPU box): Farneback engine vs the numpy CUDA-semantics model and live cv2."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, cv2
import opencv_contrib_b200 as ocb
from oracle import synth, metrics, farneback_gpu_model as fm

dev = torch.device("cuda:0")


def run(I0, I1, init=None, **kw):
    alg = ocb.FarnebackOpticalFlow_create(**kw)
    d0, d1 = torch.from_numpy(I0).to(dev), torch.from_numpy(I1).to(dev)
    fl = None if init is None else torch.from_numpy(init.copy()).to(dev)
    f = alg.calc(d0, d1, fl)
    torch.cuda.synchronize()
    return f.cpu().numpy(), alg


for (h, w, kind) in [(240, 320, "smooth"), (243, 317, "affine")]:
    I0, I1, gt = synth.make_pair(h, w, seed=1, kind=kind)
    for kw in [dict(), dict(flags=256), dict(fastPyramids=True), dict(polyN=7, polySigma=1.5, pyrScale=0.8, numLevels=3),
               dict(pyrScale=0.3, numLevels=3), dict(winSize=9, numIters=3)]:
        got, alg = run(I0, I1, **kw)
        ref = fm.calc(I0, I1, fm.FarnebackParams(**kw))
        cpu = cv2.calcOpticalFlowFarneback(I0, I1, None, kw.get("pyrScale", 0.5), kw.get("numLevels", 5), kw.get("winSize", 13),
                                           kw.get("numIters", 10), kw.get("polyN", 5), kw.get("polySigma", 1.1), kw.get("flags", 0))
        print(h, w, kind, kw, "\n   vs model:", metrics.epe_stats(got, ref), "\n   vs cv2  :", metrics.epe_stats(got, cpu),
              "ncc", metrics.ncc_dissimilarity(got, cpu), " finite", np.isfinite(got).all(), flush=True)
# initial flow
I0, I1, gt = synth.make_pair(240, 320, seed=3, kind="smooth")
init = (gt + 0.3).astype(np.float32)
got, _ = run(I0, I1, init=init, flags=4)
ref = fm.calc(I0, I1, fm.FarnebackParams(flags=4), init_flow=init)
cpu = cv2.calcOpticalFlowFarneback(I0, I1, init.copy(), 0.5, 5, 13, 10, 5, 1.1, 4)
print("init flow: vs model", metrics.epe_stats(got, ref), "vs cv2", metrics.epe_stats(got, cpu))
# f32 input
got, _ = run(I0.astype(np.float32), I1.astype(np.float32))
ref = fm.calc(I0, I1)
print("f32 input vs model", metrics.epe_stats(got, ref))

# timing 1080p
I0, I1, gt = synth.make_pair(1080, 1920, seed=0, kind="smooth")
d0, d1 = torch.from_numpy(I0).to(dev), torch.from_numpy(I1).to(dev)
for kw in [dict(), dict(fastPyramids=True), dict(flags=256)]:
    alg = ocb.FarnebackOpticalFlow_create(**kw)
    flow = torch.empty((1080, 1920, 2), dtype=torch.float32, device=dev)
    for _ in range(3):
        alg.calc(d0, d1, flow)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        alg.calc(d0, d1, flow)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print("1080p farneback", kw, ": %.3f ms/pair = %.1f pairs/s" % (ms, 1000 / ms))
    alg.setProfiling(True)
    alg.resetStats()
    alg.calc(d0, d1, flow)
    torch.cuda.synchronize()
    print("   profile:", json.dumps(alg.getStats()["classes"]))
    if not kw:
        f = flow.cpu().numpy()
        t = time.time()
        cpu = cv2.calcOpticalFlowFarneback(I0, I1, None, 0.5, 5, 13, 10, 5, 1.1, 0)
        dt = time.time() - t
        print("   1080p vs cv2:", metrics.epe_stats(f, cpu), "ncc", metrics.ncc_dissimilarity(f, cpu),
              "cv2 time %.2fs threads %d" % (dt, cv2.getNumThreads()))
