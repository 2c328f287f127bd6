"""Generates the committed golden fixtures from the live CPU reference in the BUILD container
(cv2 4.13.0: cv2.calcOpticalFlowFarneback = opencv/opencv modules/video/src/optflowgf.cpp, the
function modules/optflow/src/interfaces.cpp:154-157 forwards to).   python tests/golden/make_golden.py"""
import os
import sys

import numpy as np
import cv2

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import synth  # noqa: E402

CASES = {
    "box": (dict(), 1e-4, 0.02),
    "poly7": (dict(polyN=7, polySigma=1.5), 1e-4, 0.02),
    "gauss": (dict(flags=256), 2e-2, 0.2),
    "scale08": (dict(pyrScale=0.8, numLevels=3), 1e-4, 0.02),
}
for name, (kw, ncc_tol, epe_tol) in CASES.items():
    I0, I1, _ = synth.make_pair(144, 192, seed=11, kind="smooth")
    flow = cv2.calcOpticalFlowFarneback(I0, I1, None, kw.get("pyrScale", 0.5), kw.get("numLevels", 5), 13, 10,
                                        kw.get("polyN", 5), kw.get("polySigma", 1.1), kw.get("flags", 0))
    out = {"I0": I0, "I1": I1, "flow": flow.astype(np.float32), "ncc_tol": ncc_tol, "epe_tol": epe_tol,
           "cv2_version": cv2.__version__}
    out.update({"kw_" + k: v for k, v in kw.items()})
    np.savez_compressed(os.path.join(HERE, f"farneback_{name}.npz"), **out)
    print(name, flow.shape)
