import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import opencv_contrib_b200 as ocb
from oracle import synth, metrics, tvl1_gpu_model as gm
dev = torch.device("cuda:0")
I0, I1, gt = synth.make_pair(243, 317, seed=1, kind="smooth")
kw = dict(nscales=4, warps=4, epsilon=0.0, iterations=30)
ref = gm.calc(I0, I1, gm.TVL1Params(**kw))
outs = {}
for path, K in [(1, 0), (0, 1), (0, 5), (0, 6), (2, 6), (0, 7)]:
    alg = ocb.OpticalFlowDual_TVL1_create(**kw)
    alg.setEngineOption("kernel_path", path); alg.setEngineOption("fused_iters", K)
    f = alg.calc(torch.from_numpy(I0).to(dev), torch.from_numpy(I1).to(dev)); torch.cuda.synchronize()
    outs[(path, K)] = f.cpu().numpy()
print("unfused vs model", metrics.epe_stats(outs[(1, 0)], ref))
for k, v in outs.items():
    print(k, "bit-equal to unfused:", np.array_equal(v, outs[(1, 0)]))
I0, I1, gt = synth.make_pair(1080, 1920, seed=0, kind="smooth")
d0, d1 = torch.from_numpy(I0).to(dev), torch.from_numpy(I1).to(dev)
for (path, K, graph) in [(0, 2, 1), (0, 4, 1), (0, 6, 1), (0, 8, 1), (0, 10, 1), (2, 6, 1), (0, 6, 0)]:
    alg = ocb.OpticalFlowDual_TVL1_create(nscales=5, warps=10, epsilon=0.0, iterations=30)
    alg.setEngineOption("kernel_path", path); alg.setEngineOption("fused_iters", K); alg.setEngineOption("use_graph", graph)
    flow = torch.empty((1080, 1920, 2), dtype=torch.float32, device=dev)
    for _ in range(2): alg.calc(d0, d1, flow)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    e0.record()
    for _ in range(n): alg.calc(d0, d1, flow)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print("1080p 5x10x30 path=%d K=%d graph=%d: %.2f ms/pair = %.1f pairs/s" % (path, K, graph, ms, 1000 / ms), flush=True)
