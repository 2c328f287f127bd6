"""Farneback 1080p defaults: aux_path sweep (registers / occupancy of the fused iteration kernel), single stream
(CUDA graph, real stream) and 8 streams, bit-equality against the default."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import opencv_contrib_b200 as ocb
from oracle import synth
dev = torch.device("cuda:0")
I0, I1, gt = synth.make_pair(1080, 1920, seed=0, kind="smooth")
d0, d1 = torch.from_numpy(I0).to(dev), torch.from_numpy(I1).to(dev)
ref = None
NS = 8
for aux in [int(a) for a in (sys.argv[1:] or ["0", "3", "4"])]:
    algs = [ocb.FarnebackOpticalFlow_create() for _ in range(NS)]
    for a in algs: a.setEngineOption("aux_path", aux)
    flows = [torch.empty((1080, 1920, 2), dtype=torch.float32, device=dev) for _ in range(NS)]
    streams = [torch.cuda.Stream() for _ in range(NS)]
    s0 = streams[0]
    for _ in range(3): algs[0].calc(d0, d1, flows[0], s0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record(s0)
    for _ in range(n): algs[0].calc(d0, d1, flows[0], s0)
    e1.record(s0); torch.cuda.synchronize()
    ms1 = e0.elapsed_time(e1) / n
    out = flows[0].cpu().numpy()
    if ref is None: ref = out
    for i in range(NS): algs[i].calc(d0, d1, flows[i], streams[i])
    torch.cuda.synchronize()
    e0.record()
    R = 6
    for r in range(R):
        for i in range(NS):
            if r == 0: streams[i].wait_event(e0)
            algs[i].calc(d0, d1, flows[i], streams[i])
    for s in streams: torch.cuda.current_stream().wait_stream(s)
    e1.record(); torch.cuda.synchronize()
    msN = e0.elapsed_time(e1) / (R * NS)
    print("aux=%d: 1 stream %.3f ms/pair (%.1f/s); %d streams %.3f ms/pair (%.1f/s); max |diff| vs default %.2e"
          % (aux, ms1, 1000 / ms1, NS, msN, 1000 / msN, float(np.abs(out - ref).max())), flush=True)
