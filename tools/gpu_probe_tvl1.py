"""First-contact probe (GPU box): prints how far each TV-L1 kernel path is from the oracles."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import opencv_contrib_b200 as ocb
from oracle import synth, metrics, tvl1_gpu_model as gm, tvl1_cpu

dev = torch.device("cuda:0")
print(torch.cuda.get_device_name(0))

def run(I0, I1, path, fused=0, graph=1, **kw):
    alg = ocb.OpticalFlowDual_TVL1_create(**kw)
    alg.setEngineOption("kernel_path", path)
    alg.setEngineOption("fused_iters", fused)
    alg.setEngineOption("use_graph", graph)
    d0, d1 = torch.from_numpy(I0).to(dev), torch.from_numpy(I1).to(dev)
    f = alg.calc(d0, d1)
    torch.cuda.synchronize()
    return f.cpu().numpy(), alg

for (h, w, kind) in [(120, 160, "const"), (243, 317, "smooth")]:
    I0, I1, gt = synth.make_pair(h, w, seed=1, kind=kind)
    kw = dict(nscales=4, warps=4, epsilon=0.0, iterations=30)
    ref = gm.calc(I0, I1, gm.TVL1Params(**kw))
    a, _ = run(I0, I1, 1, **kw)
    print(h, w, kind, "unfused vs model:", metrics.epe_stats(a, ref))
    for K in (1, 2, 3, 5, 6, 10):
        for graph in (0, 1):
            b, alg = run(I0, I1, 0, fused=K, graph=graph, **kw)
            print("  blocked K=%d graph=%d vs unfused: maxabs=%g  bit-equal=%s  launches=%d" % (
                K, graph, np.abs(a - b).max(), np.array_equal(a, b), alg.getStats()["launches"]))
    cpu = tvl1_cpu.calc(I0, I1, tvl1_cpu.TVL1Params(nscales=4, warps=4, epsilon=0.0, innerIterations=1,
                                                    outerIterations=30, medianFiltering=1))
    print("  gpu vs cpu-oracle (interior 16):", metrics.epe_stats(a, cpu, border=16), "ncc", metrics.ncc_dissimilarity(a[16:-16,16:-16], cpu[16:-16,16:-16]))
    print("  gpu vs gt (interior 16):", metrics.epe_stats(a, gt, border=16))

# epsilon > 0 cadence
I0, I1, gt = synth.make_pair(120, 160, seed=2, kind="const")
kw = dict(nscales=3, warps=3, epsilon=0.01, iterations=100)
tr = []
ref = gm.calc(I0, I1, gm.TVL1Params(**kw), trace=tr)
a, alg = run(I0, I1, 0, **kw)
print("eps>0: vs model", metrics.epe_stats(a, ref), "model iters", tr, "engine iters", alg.getStats()["iterations_run"])
# gamma
kw = dict(nscales=3, warps=3, epsilon=0.0, iterations=20, gamma=1.0)
ref = gm.calc(I0, I1, gm.TVL1Params(**kw))
a, alg = run(I0, I1, 0, **kw)
print("gamma=1: vs model", metrics.epe_stats(a, ref))
# f32 input
I0f, I1f, _ = synth.make_pair(120, 160, seed=2, kind="const", dtype="f32")
kw = dict(nscales=3, warps=3, epsilon=0.0, iterations=20)
ref = gm.calc(I0f, I1f, gm.TVL1Params(**kw))
a, alg = run(I0f, I1f, 0, **kw)
print("f32 in: vs model", metrics.epe_stats(a, ref))

# timing 1080p
I0, I1, gt = synth.make_pair(1080, 1920, seed=0, kind="smooth")
d0, d1 = torch.from_numpy(I0).to(dev), torch.from_numpy(I1).to(dev)
for (path, K, graph) in [(1, 0, 0), (0, 1, 1), (0, 2, 1), (0, 3, 1), (0, 5, 1), (0, 6, 1), (0, 10, 1), (0, 5, 0)]:
    alg = ocb.OpticalFlowDual_TVL1_create(nscales=5, warps=10, epsilon=0.0, iterations=30)
    alg.setEngineOption("kernel_path", path); alg.setEngineOption("fused_iters", K); alg.setEngineOption("use_graph", graph)
    flow = torch.empty((1080, 1920, 2), dtype=torch.float32, device=dev)
    for _ in range(2): alg.calc(d0, d1, flow)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 3
    e0.record()
    for _ in range(n): alg.calc(d0, d1, flow)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print("1080p 5x10x30 path=%d K=%d graph=%d: %.2f ms/pair = %.1f pairs/s" % (path, K, graph, ms, 1000 / ms))
    if path == 0 and K == 5 and graph == 1:
        f = flow.cpu().numpy()
        print("   1080p gpu vs gt interior:", metrics.epe_stats(f, gt, border=32))
    if graph == 0 or path == 1:
        alg.setProfiling(True); alg.resetStats(); alg.calc(d0, d1, flow); torch.cuda.synchronize()
        print("   profile:", json.dumps(alg.getStats()["classes"]))
