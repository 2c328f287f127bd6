"""Timing probe: Brox class breakdown, DensePyrLK, video front end throughput, interpolateFrames."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import opencv_contrib_b200 as ocb
from oracle import synth, metrics
dev = torch.device("cuda:0")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timeit(fn, n, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


what = sys.argv[1:] or ["brox", "lk", "video", "interp"]
if "brox" in what:
    I0, I1, gt = synth.make_pair(720, 1280, seed=0, kind="smooth", dtype="f32")
    d0, d1 = torch.from_numpy(I0).to(dev), torch.from_numpy(I1).to(dev)
    flow = torch.empty((720, 1280, 2), dtype=torch.float32, device=dev)
    for path in (0, 2, 1):
        alg = ocb.BroxOpticalFlow_create(0.197, 50.0, 0.8, 10, 77, 10)
        alg.setEngineOption("kernel_path", path)
        ms = timeit(lambda: alg.calc(d0, d1, flow), 5)
        print("brox 720p (10,77,10) path=%d: %.2f ms/pair = %.1f pairs/s" % (path, ms, 1000 / ms), "vs gt",
              metrics.epe_stats(flow.cpu().numpy(), gt, border=32)["mean"], flush=True)
        alg.setProfiling(True); alg.resetStats(); alg.calc(d0, d1, flow); torch.cuda.synchronize()
        print("   ", {k: (round(v["ms"], 3), v["launches"]) for k, v in alg.getStats()["classes"].items()}, flush=True)
if "broxlevels" in what:
    # per-level cost: solve at 720p * 0.8^k and difference consecutive totals
    prev = None
    for k in range(0, 10):
        sc = 0.8 ** k
        h, w = int(np.ceil(720 * sc)), int(np.ceil(1280 * sc))
        I0, I1, gt = synth.make_pair(h, w, seed=0, kind="smooth", dtype="f32")
        d0, d1 = torch.from_numpy(I0).to(dev), torch.from_numpy(I1).to(dev)
        flow = torch.empty((h, w, 2), dtype=torch.float32, device=dev)
        alg = ocb.BroxOpticalFlow_create(0.197, 50.0, 0.8, 10, 77, 10)
        ms = timeit(lambda: alg.calc(d0, d1, flow), 5)
        alg.setProfiling(True); alg.resetStats(); alg.calc(d0, d1, flow); torch.cuda.synchronize()
        cl = {kk: (round(v["ms"], 3), v["launches"]) for kk, v in alg.getStats()["classes"].items()}
        print("brox %4dx%-4d levels %2d: %.3f ms" % (w, h, alg.getStats()["levels"], ms), cl, flush=True)
if "lk" in what:
    I0, I1, gt = synth.make_pair(1080, 1920, seed=0, kind="smooth")
    d0, d1 = torch.from_numpy(I0).to(dev), torch.from_numpy(I1).to(dev)
    flow = torch.zeros((1080, 1920, 2), dtype=torch.float32, device=dev)
    for path in (0, 1):
        alg = ocb.DensePyrLKOpticalFlow_create()
        alg.setEngineOption("kernel_path", path)
        ms = timeit(lambda: alg.calc(d0, d1, flow), 3)
        print("denselk 1080p (13x13,3,30) path=%d: %.2f ms/pair = %.1f pairs/s" % (path, ms, 1000 / ms), "vs gt",
              metrics.epe_stats(flow.cpu().numpy(), gt, border=32)["mean"], flush=True)
if "video" in what:
    import cv2
    T = synth.texture(1080 + 64, 1920 + 64, 11)
    frames = []
    for k in range(8):
        M = np.float32([[1, 0, -32 + 1.5 * k], [0, 1, -32 - 0.75 * k]])
        frames.append(np.clip(cv2.warpAffine(T, M, (1920, 1080), flags=cv2.INTER_LINEAR | cv2.WARP_INVERSE_MAP), 0, 255).astype(np.uint8))
    for name, make in [("tvl1 5x10x30", lambda: ocb.OpticalFlowDual_TVL1_create(0.25, 0.15, 0.3, 5, 10, 0.0, 30)),
                       ("tvl1 defaults", lambda: ocb.OpticalFlowDual_TVL1_create()),
                       ("farneback", lambda: ocb.FarnebackOpticalFlow_create())]:
        for warm in (False, True):
            vf = ocb.VideoFlow(make(), 1080, 1920, depth=3, warm_start=warm)
            seq = [frames[i % 8] for i in range(41)]
            for _ in vf.run(seq[:5], copy=False): pass
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            n = 0
            last = None
            for p, fl in vf.run(seq, copy=False):
                n += 1; last = float(np.abs(fl[::8, ::8, 0]).mean())
            dt = time.perf_counter() - t0
            print("video %-14s warm=%d: %.1f pairs/s (host frames in, host flows out, %d pairs), mean |u| %.3f" % (name, warm, n / dt, n, last), flush=True)
            vf.close()
if "interp" in what:
    f0, f1 = torch.rand((1080, 1920), device=dev), torch.rand((1080, 1920), device=dev)
    fu, fv = torch.randn((1080, 1920), device=dev) * 3, torch.randn((1080, 1920), device=dev) * 3
    out, buf = torch.empty_like(f0), torch.empty((6 * 1080, 1920), device=dev)
    for corr in (False, True):
        ms = timeit(lambda: ocb.interpolateFrames(f0, f1, fu, fv, -fu, -fv, 0.5, out, buf, corrected=corr), 20)
        print("interpolateFrames 1080p corrected=%d: %.3f ms (%.0f frames/s)" % (corr, ms, 1000 / ms), flush=True)
