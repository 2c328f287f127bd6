"""Full-size parity numbers for DESIGN.md: TV-L1 BASELINE config vs the CPU oracle, Farneback vs cv2."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, cv2
import opencv_contrib_b200 as ocb
from oracle import synth, metrics, tvl1_cpu, tvl1_cpu_native
dev = torch.device("cuda:0")
for kind in ("smooth", "affine"):
    I0, I1, gt = synth.make_pair(1080, 1920, seed=0, kind=kind)
    d0, d1 = torch.from_numpy(I0).to(dev), torch.from_numpy(I1).to(dev)
    got = ocb.OpticalFlowDual_TVL1_create(nscales=5, warps=10, epsilon=0.0, iterations=30).calc(d0, d1).cpu().numpy()
    t0 = time.time()
    cpu = tvl1_cpu_native.calc(I0, I1, tvl1_cpu.TVL1Params(nscales=5, warps=10, epsilon=0.0, innerIterations=1, outerIterations=30, medianFiltering=1))
    dt = time.time() - t0
    st = metrics.epe_stats(got, cpu, border=32)
    print("tvl1 1080p 5x10x30", kind, "vs CPU oracle:", {k: round(float(v), 5) for k, v in st.items()},
          "ncc %.2e" % metrics.ncc_dissimilarity(got[32:-32, 32:-32], cpu[32:-32, 32:-32]),
          "| vs truth gpu %.4f cpu %.4f | cpu %.1f s" % (metrics.epe_stats(got, gt, border=32)["mean"], metrics.epe_stats(cpu, gt, border=32)["mean"], dt), flush=True)
    f = ocb.FarnebackOpticalFlow_create().calc(d0, d1).cpu().numpy()
    c = cv2.calcOpticalFlowFarneback(I0, I1, None, 0.5, 5, 13, 10, 5, 1.1, 0)
    print("farneback 1080p", kind, "vs cv2:", {k: round(float(v), 7) for k, v in metrics.epe_stats(f, c).items()}, "ncc %.2e" % metrics.ncc_dissimilarity(f, c), flush=True)
