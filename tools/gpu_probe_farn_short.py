import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, cv2
import opencv_contrib_b200 as ocb
from oracle import synth, metrics
dev = torch.device("cuda:0")
I0, I1, gt = synth.make_pair(1080, 1920, seed=0, kind="smooth")
d0, d1 = torch.from_numpy(I0).to(dev), torch.from_numpy(I1).to(dev)
cpu = cv2.calcOpticalFlowFarneback(I0, I1, None, 0.5, 5, 13, 10, 5, 1.1, 0)
for path in (3, 0):
    alg = ocb.FarnebackOpticalFlow_create()
    alg.setEngineOption("kernel_path", path)
    flow = torch.empty((1080, 1920, 2), dtype=torch.float32, device=dev)
    for _ in range(3): alg.calc(d0, d1, flow)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n): alg.calc(d0, d1, flow)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    f = flow.cpu().numpy()
    alg.setProfiling(True); alg.resetStats(); alg.calc(d0, d1, flow); torch.cuda.synchronize()
    cl = alg.getStats()["classes"]
    print("farneback 1080p path=%d: %.3f ms/pair = %.1f pairs/s; vs cv2 ncc %.2e mean epe %.2e; iter %.3f ms" % (
        path, ms, 1000 / ms, metrics.ncc_dissimilarity(f, cpu), metrics.epe_stats(f, cpu)["mean"], cl["farn_iter"]["ms"]), flush=True)
    print("   ", {k: round(v["ms"], 3) for k, v in cl.items()})
