"""The C++ adapter (include/b200flow/cudaoptflow_compat.hpp) used from a C++ call site must give
bit-identical flow to the ctypes path (same library, same kernels)."""
import os
import subprocess

import numpy as np
import pytest

from oracle import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_compat")


@pytest.mark.parametrize("algo", ["tvl1", "farneback"])
def test_cpp_call_site_matches_ctypes_path(cuda_device, tmp_path, algo):
    import torch
    import opencv_contrib_b200 as ocb
    if not os.path.exists(EXE):
        subprocess.run(["make", "-C", os.path.dirname(EXE)], check=True)
    I0, I1, _ = synth.make_pair(120, 168, seed=12, kind="smooth")
    p0, p1, po, pm = tmp_path / "i0.raw", tmp_path / "i1.raw", tmp_path / "flow.raw", tmp_path / "mid.raw"
    I0.tofile(p0)
    I1.tofile(p1)
    r = subprocess.run([EXE, algo, "120", "168", str(p0), str(p1), str(po), str(pm)], capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    got = np.fromfile(po, np.float32).reshape(120, 168, 2)
    if algo == "tvl1":
        alg = ocb.OpticalFlowDual_TVL1_create(nscales=3, warps=2, epsilon=0.0, iterations=20)
    else:
        alg = ocb.FarnebackOpticalFlow_create()
    ref = alg.calc(torch.from_numpy(I0).to(cuda_device), torch.from_numpy(I1).to(cuda_device)).cpu().numpy()
    assert np.array_equal(got, ref)
    # the C++ interpolateFrames call site vs the ctypes one on the same inputs (float atomics: rounding-level)
    dev = cuda_device
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    f0, f1 = (I0 / np.float32(255)).astype(np.float32), (I1 / np.float32(255)).astype(np.float32)
    mid = ocb.interpolateFrames(t(f0), t(f1), t(ref[..., 0]), t(ref[..., 1]), t(-ref[..., 0]), t(-ref[..., 1]), 0.5,
                                corrected=True)
    got_mid = np.fromfile(pm, np.float32).reshape(120, 168)
    # float atomics: wherever the splat coverage is tiny the normalisation amplifies the order-dependent rounding,
    # so two runs of the same call agree almost everywhere rather than everywhere
    d = np.abs(got_mid - mid.cpu().numpy())
    assert np.isfinite(got_mid).all() and float((d < 1e-4).mean()) > 0.995, float((d < 1e-4).mean())
