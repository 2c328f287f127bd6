"""Where a pass of the TV-L1 iteration kernel spends its time: level-0-only solves (nscales = 1) on images that are an
exact number of waves of 48 x 48 tiles over 148 SMs, with the iteration count of the pass overridden (B2F_DBG_TVL1_ITERS)
so that  T(waves, iters) = launch + waves * (tile_fixed + iters * per_iteration)  can be solved for its three terms.
Also times the warp kernel variants (aux_path 0 separable / 3 tiled) on the same images and on 1080p."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import opencv_contrib_b200 as ocb
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)


def classes(h, w, aux=0, iters_dbg=None, reps=7):
    I0 = rng.integers(0, 255, (h, w), dtype=np.uint8)
    I1 = np.roll(I0, 2, axis=1)
    d0, d1 = torch.from_numpy(I0).to(dev), torch.from_numpy(I1).to(dev)
    alg = ocb.OpticalFlowDual_TVL1_create(nscales=1, warps=1, epsilon=0.0, iterations=8)
    alg.setEngineOption("aux_path", aux)
    alg.setEngineOption("use_graph", 0)
    flow = torch.empty((h, w, 2), dtype=torch.float32, device=dev)
    if iters_dbg is not None: os.environ["B2F_DBG_TVL1_ITERS"] = str(iters_dbg)
    else: os.environ.pop("B2F_DBG_TVL1_ITERS", None)
    out = {}
    for r in range(reps + 2):
        alg.setProfiling(True); alg.resetStats(); alg.calc(d0, d1, flow); torch.cuda.synchronize()
        st = alg.getStats()
        if r >= 2:
            for k, v in st["classes"].items():
                if v["launches"]: out.setdefault(k, []).append(1000.0 * v["ms"] / v["launches"])
    os.environ.pop("B2F_DBG_TVL1_ITERS", None)
    return {k: float(np.median(v)) for k, v in out.items()}


W = 48 * 148
for waves in (() if "warp" in sys.argv[1:] else (1, 2, 4, 8)):
    row = []
    for it in (0, 4, 8):
        c = classes(48 * waves, W, iters_dbg=it)
        row.append(c["tvl1_iter"])
    print("waves %d: iter-kernel launch us at 0 / 4 / 8 iterations: %.2f %.2f %.2f  -> per iteration per wave %.3f us, fixed %.2f us"
          % (waves, row[0], row[1], row[2], (row[2] - row[0]) / 8 / waves, row[0]), flush=True)
for (h, w) in ((384, W), (2160, 3840), (1080, 1920), (864, 1536), (691, 1229), (553, 983), (442, 786)):
    a = classes(h, w, aux=0)["tvl1_warp"]
    b = classes(h, w, aux=3)["tvl1_warp"]
    print("warp kernel %dx%d: separable %.2f us, tiled %.2f us" % (w, h, a, b), flush=True)
