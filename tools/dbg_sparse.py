import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2, numpy as np, torch
import opencv_contrib_b200 as ocb
from oracle import synth
dev = torch.device("cuda:0")
for kind in ("const", "affine"):
    I0, I1, gt = synth.make_pair(360, 480, seed=21, kind=kind)
    pts = cv2.goodFeaturesToTrack(I0, 1000, 0.01, 0.0).reshape(-1, 2).astype(np.float32)
    gold, stg, _ = cv2.calcOpticalFlowPyrLK(I0, I1, pts.reshape(-1, 1, 2), None)
    gold, stg = gold.reshape(-1, 2), stg.ravel()
    alg = ocb.SparsePyrLKOpticalFlow_create()
    nxt, st, err = alg.calc(torch.from_numpy(I0).to(dev), torch.from_numpy(I1).to(dev), torch.from_numpy(pts).to(dev), wantErr=True)
    nxt, st = nxt.cpu().numpy(), st.cpu().numpy()
    g = gt[np.clip(pts[:, 1].astype(int), 0, 359), np.clip(pts[:, 0].astype(int), 0, 479)]
    n_st = 0
    for i in range(len(pts)):
        a, b = nxt[i].astype(np.int32), gold[i].astype(np.int32)
        bad_st = bool(st[i]) != bool(stg[i])
        bad_xy = (not bad_st) and st[i] and (abs(a[0] - b[0]) > 1 or abs(a[1] - b[1]) > 1)
        if bad_st or bad_xy:
            print(kind, i, "pt", pts[i], "gpu", nxt[i], int(st[i]), "cpu", gold[i], int(stg[i]), "true", pts[i] + g[i], "status-mismatch" if bad_st else "xy")
