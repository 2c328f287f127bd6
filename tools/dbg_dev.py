import sys, os, threading, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import opencv_contrib_b200 as ocb
from oracle import synth
d1 = torch.device("cuda:1")
I0, I1, _ = synth.make_pair(96, 128, seed=2, kind="const")
ref = ocb.FarnebackOpticalFlow_create(numLevels=3).calc_host(I0, I1)
def work():
    try:
        print("thread device", torch.cuda.current_device(), flush=True)
        s = torch.cuda.Stream(device=d1)
        print("stream", hex(s.cuda_stream), "device after stream create", torch.cuda.current_device(), flush=True)
        alg = ocb.FarnebackOpticalFlow_create(numLevels=3)
        h = alg.calc_host(I0, I1, None, s)
        print("host ok", np.array_equal(h, ref), torch.cuda.current_device(), flush=True)
        a, b = torch.from_numpy(I0).to(d1), torch.from_numpy(I1).to(d1)
        f = alg.calc(a, b, torch.empty((96, 128, 2), device=d1), s)
        s.synchronize()
        print("dev ok", np.array_equal(f.cpu().numpy(), ref), torch.cuda.current_device(), flush=True)
        alg2 = ocb.OpticalFlowDual_TVL1_create(nscales=3, warps=2, epsilon=0.0, iterations=16)
        f2 = alg2.calc(a, b, torch.empty((96, 128, 2), device=d1), s); s.synchronize()
        r2 = ocb.OpticalFlowDual_TVL1_create(nscales=3, warps=2, epsilon=0.0, iterations=16).calc_host(I0, I1)
        print("tvl1 dev1 vs dev0", np.array_equal(f2.cpu().numpy(), r2), flush=True)
    except BaseException:
        traceback.print_exc()
t = threading.Thread(target=work); t.start(); t.join()
