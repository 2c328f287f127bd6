"""CPU tests for the host-side adjacent components: Middlebury .flo files and the
reference's error measures (csrc/flowio.cu through the C ABI).  cv2.readOpticalFlow / writeOpticalFlow are
the live reference for the file format; the statistics are checked against direct numpy restatements of
optflow/samples/optical_flow_evaluation.cpp:23-163 and optflow/test/test_tvl1optflow.cpp:114-142.
Also pins the numpy interpolateFrames model on cases with a closed-form answer.
"""
import os
import struct

import cv2
import numpy as np
import pytest

from opencv_contrib_b200 import flowio
from opencv_contrib_b200._lib import B2FError
from oracle import interpolate_model as im


def _rand_flow(h, w, seed=0):
    return np.random.default_rng(seed).normal(0, 3, (h, w, 2)).astype(np.float32)


def test_flo_roundtrip_and_layout(tmp_path):
    f = _rand_flow(37, 53)
    p = str(tmp_path / "a.flo")
    flowio.writeOpticalFlow(p, f)
    raw = open(p, "rb").read()
    assert raw[:4] == b"PIEH" and struct.unpack("<f", raw[:4])[0] == 202021.25
    assert struct.unpack("<ii", raw[4:12]) == (53, 37)
    assert len(raw) == 12 + 37 * 53 * 8
    assert np.array_equal(flowio.readOpticalFlow(p), f)
    # non-contiguous rows (ROI of a wider array) are written row by row
    big = np.zeros((37, 80, 2), np.float32)
    big[:, 5:58] = f
    flowio.writeOpticalFlow(p, big[:, 5:58])
    assert np.array_equal(flowio.readOpticalFlow(p), f)


def test_flo_interoperates_with_cv2(tmp_path):
    f = _rand_flow(21, 64, seed=2)
    p1, p2 = str(tmp_path / "ours.flo"), str(tmp_path / "cv.flo")
    flowio.writeOpticalFlow(p1, f)
    assert np.array_equal(cv2.readOpticalFlow(p1), f)
    assert cv2.writeOpticalFlow(p2, f)
    assert np.array_equal(flowio.readOpticalFlow(p2), f)
    assert open(p1, "rb").read() == open(p2, "rb").read()


def test_flo_errors(tmp_path):
    p = str(tmp_path / "bad.flo")
    open(p, "wb").write(b"NOPE" + b"\0" * 64)
    with pytest.raises(B2FError):
        flowio.readOpticalFlow(p)  # wrong tag (reference: CV_Assert(tag == FLO_TAG_FLOAT))
    with pytest.raises(B2FError):
        flowio.readOpticalFlow(str(tmp_path / "missing.flo"))
    open(p, "wb").write(b"PIEH" + struct.pack("<ii", 8, 8) + b"\0" * 100)  # truncated payload
    with pytest.raises(B2FError):
        flowio.readOpticalFlow(p)


def test_error_maps_follow_the_sample():
    a, b = _rand_flow(40, 50, 1), _rand_flow(40, 50, 2)
    a[3, 4, 0] = np.nan
    b[7, 8, 1] = 2e9  # |component| >= 1e9 counts as invalid
    e = flowio.errorMap(a, b, flowio.ERR_ENDPOINT)
    assert np.isnan(e[3, 4]) and np.isnan(e[7, 8]) and np.isnan(e).sum() == 2
    ok = ~np.isnan(e)
    d = a - b
    ref = np.sqrt((d[..., 0].astype(np.float64) ** 2 + d[..., 1].astype(np.float64) ** 2).astype(np.float32))
    assert np.array_equal(e[ok], ref[ok])
    # angular: the sample's precedence (dot / |u1| * |u2|) vs the intended formula
    a2, b2 = 0.05 * a, 0.05 * b  # keeps the sample's argument inside acos' domain for most pixels
    er = flowio.errorMap(a2, b2, flowio.ERR_ANGULAR_REFERENCE)
    ei = flowio.errorMap(a2, b2, flowio.ERR_ANGULAR)
    dot = (a2[..., 0].astype(np.float64) * b2[..., 0] + a2[..., 1].astype(np.float64) * b2[..., 1] + 1.0)
    n1 = np.sqrt(a2[..., 0].astype(np.float64) ** 2 + a2[..., 1].astype(np.float64) ** 2 + 1)
    n2 = np.sqrt(b2[..., 0].astype(np.float64) ** 2 + b2[..., 1].astype(np.float64) ** 2 + 1)
    with np.errstate(invalid="ignore"):
        ref_r = np.arccos((dot / n1 * n2).astype(np.float32))
        ref_i = np.arccos((dot / (n1 * n2)).astype(np.float32))
    m = ok & ~np.isnan(ref_r)
    assert np.allclose(er[m], ref_r[m], atol=1e-6) and np.allclose(ei[ok], ref_i[ok], atol=1e-6)
    assert np.array_equal(np.isnan(er[ok]), np.isnan(ref_r[ok]))


def test_error_stats_follow_the_sample():
    rng = np.random.default_rng(5)
    err = np.abs(rng.normal(0, 2.5, (64, 80))).astype(np.float32)
    mask = (rng.random((64, 80)) > 0.3).astype(np.uint8)
    for m in (None, mask):
        s = flowio.errorStats(err, m)
        sel = err[m != 0] if m is not None else err.ravel()
        mean, std = cv2.meanStdDev(err, mask=m)
        assert s["count"] == sel.size
        assert abs(s["mean"] - float(mean[0, 0])) < 1e-9 and abs(s["std"] - float(std[0, 0])) < 1e-7
        for thr, r in s["R"].items():
            assert abs(r - np.float32((sel > thr).sum()) / sel.size) < 1e-7
        # A statistics: the reference's 1024-bin calcHist walk
        mx = float(sel.max())
        hist = cv2.calcHist([err], [0], m, [1024], [0, mx]).ravel()
        for q, a in s["A"].items():
            cutoff = int(np.floor(np.float32(q) * sel.size + np.float32(0.5)))
            counter, b = 0, 0
            while b < 1024 and counter < cutoff:
                counter += int(hist[b])
                b += 1
            assert abs(a - np.float32(b) / 1024 * np.float32(mx)) < 1e-6, (q, a)


def test_accuracy_criterion():
    g = _rand_flow(30, 30, 3)
    f = g.copy()
    f[:3] += 0.2      # 10 % of the rows are off by more than the threshold
    g[5, 5] = np.nan   # invalid gold pixels are not counted
    f[6, 6] = np.nan   # invalid result on a valid gold pixel counts as a miss
    frac = flowio.accuracy(g, f, 0.1)
    assert abs(frac - (900 - 90 - 1 - 1) / 899) < 1e-12


# --------------------------------------------------------------------------- interpolateFrames model
def test_interpolate_model_zero_flow_is_identity():
    rng = np.random.default_rng(0)
    f0, f1 = rng.random((20, 28), dtype=np.float32), rng.random((20, 28), dtype=np.float32)
    z = np.zeros_like(f0)
    for corrected in (False, True):
        out, buf = im.interpolate_frames(f0, f1, z, z, z, z, 0.25, corrected=corrected)
        # zero flow: every pixel splats onto itself with weight 0 (dx = dy = 0 -> the (1-dx)(1-dy) tap at
        # (x, y) carries weight 1), both frames visible everywhere
        assert np.allclose(buf[0, :, :28], 1) and np.allclose(buf[1, :, :28], 1)
        want = 0.75 * f0 + 0.25 * (f1 if corrected else f0)
        assert np.allclose(out, want, atol=1e-6)


def test_interpolate_model_constant_flow_corrected():
    # frame1(x) = frame0(x - 2): forward flow (+2, 0), backward (-2, 0); halfway frame = frame0(x - 1)
    rng = np.random.default_rng(1)
    base = rng.random((16, 64), dtype=np.float32)
    f0, f1 = base, np.roll(base, 2, axis=1)
    fu, bu = np.full_like(f0, 2), np.full_like(f0, -2)
    z = np.zeros_like(f0)
    out, _ = im.interpolate_frames(f0, f1, fu, z, bu, z, 0.5, corrected=True)
    assert np.allclose(out[:, 4:-4], np.roll(base, 1, axis=1)[:, 4:-4], atol=1e-6)


def test_interpolate_model_reference_defects_are_visible():
    rng = np.random.default_rng(2)
    f0, f1 = rng.random((12, 20), dtype=np.float32), rng.random((12, 20), dtype=np.float32)
    fu = np.full_like(f0, 1.5); fv = np.full_like(f0, 0.5)
    bu = np.full_like(f0, -1.5); bv = np.full_like(f0, -0.5)
    _, ref = im.interpolate_frames(f0, f1, fu, fv, bu, bv, 0.5, stride=24, corrected=False)
    _, fix = im.interpolate_frames(f0, f1, fu, fv, bu, bv, 0.5, stride=24, corrected=True)
    assert np.all(ref[5] == 0) and np.any(fix[5] != 0)          # bwdV never written
    cleared_rows = (20 * 12) // 24                               # MemsetKernel clears the first w*h floats
    assert np.allclose(ref[0, cleared_rows + 1:, :20][fix[0, cleared_rows + 1:, :20] > 0]
                       / fix[0, cleared_rows + 1:, :20][fix[0, cleared_rows + 1:, :20] > 0], 2, atol=1e-5)


# --------------------------------------------------------------------------- property tests (hypothesis)
from hypothesis import given, settings, strategies as st  # noqa: E402


@settings(max_examples=25, deadline=None)
@given(h=st.integers(1, 40), w=st.integers(1, 40), pad=st.integers(0, 7), seed=st.integers(0, 2 ** 16))
def test_flo_roundtrip_any_shape_and_pitch(tmp_path_factory, h, w, pad, seed):
    rng = np.random.default_rng(seed)
    big = rng.normal(0, 5, (h, w + pad, 2)).astype(np.float32)
    f = big[:, pad // 2:pad // 2 + w]                    # row pitch != width * 8 when pad > 0
    p = str(tmp_path_factory.mktemp("flo") / "x.flo")
    flowio.writeOpticalFlow(p, f)
    g = flowio.readOpticalFlow(p)
    assert g.shape == (h, w, 2) and np.array_equal(g, f)
    assert np.array_equal(cv2.readOpticalFlow(p), f)    # byte-compatible with the live reference reader


@settings(max_examples=25, deadline=None)
@given(h=st.integers(2, 30), w=st.integers(2, 30), seed=st.integers(0, 2 ** 16), shift=st.floats(-3, 3))
def test_error_measures_invariants(h, w, seed, shift):
    rng = np.random.default_rng(seed)
    a = rng.normal(0, 2, (h, w, 2)).astype(np.float32)
    b = (a + np.float32(shift)).astype(np.float32)
    e = flowio.errorMap(a, b)
    assert e.shape == (h, w) and (e >= 0).all()
    assert np.allclose(e, np.abs(np.float32(shift)) * np.sqrt(2), atol=1e-5)
    assert np.array_equal(flowio.errorMap(b, a), e)                      # symmetric
    assert (flowio.errorMap(a, a) == 0).all()                            # identity
    s = flowio.errorStats(e)
    assert s["count"] == h * w and abs(s["mean"] - float(e.astype(np.float64).mean())) < 1e-9
    assert all(0.0 <= r <= 1.0 for r in s["R"].values())
    assert s["R"][0.5] >= s["R"][1.0] >= s["R"][2.0] >= s["R"][5.0] >= s["R"][10.0]   # monotone in the threshold
    assert s["A"][0.5] <= s["A"][0.75] + 1e-9 <= s["A"][0.95] + 2e-9                  # monotone in the quantile
    assert flowio.accuracy(a, a, 0.1) == 1.0
    ang = flowio.errorMap(a, a, flowio.ERR_ANGULAR)
    assert np.nanmax(np.abs(ang)) < 1e-3                                 # acos(1) up to float rounding
