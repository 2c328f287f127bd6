"""numpy model of the reference's dense pyramidal Lucas-Kanade, ``cv::cuda::DensePyrLKOpticalFlow``
(modules/cudaoptflow/src/pyrlk.cpp:238-299 + src/cuda/pyrlk.cu:709-855, pyramid by cuda::pyrDown).

TEST INFRASTRUCTURE (see oracle/__init__.py).  **Parity unpinned** even upstream: the reference has
no accuracy test for the dense path (perf_optflow.cpp:219-220 uses SANITY_CHECK_NOTHING) and no CPU
implementation; this restatement carries the reference's quirks (int-truncated patch, int32
accumulators with wrap-around, no write for rejected pixels) and the engine is checked against it.
The hardware bilinear filter of the reference is modelled with 8-bit fractional weights (CUDA
programming guide); rounding inside the texture unit cannot be reproduced bit for bit.
"""
from __future__ import annotations

import numpy as np
import cv2

F = np.float32
I32 = np.int32


def _wrap32(a):
    """int64 -> int32 two's-complement wrap-around."""
    return ((a + 2 ** 31) % 2 ** 32 - 2 ** 31).astype(np.int64)


def _texel(P, y, x):
    h, w = P.shape
    return P[np.clip(y, 0, h - 1), np.clip(x, 0, w - 1)]


def _bilinear(P, y, x):
    xb = (x - F(0.5)).astype(F)
    yb = (y - F(0.5)).astype(F)
    fx, fy = np.floor(xb), np.floor(yb)
    ax = (np.floor((xb - fx) * F(256) + F(0.5)) * F(1 / 256)).astype(F)
    ay = (np.floor((yb - fy) * F(256) + F(0.5)) * F(1 / 256)).astype(F)
    x0, y0 = fx.astype(np.int64), fy.astype(np.int64)
    t00, t01 = _texel(P, y0, x0), _texel(P, y0, x0 + 1)
    t10, t11 = _texel(P, y0 + 1, x0), _texel(P, y0 + 1, x0 + 1)
    return ((F(1) - ax) * (F(1) - ay) * t00 + ax * (F(1) - ay) * t01 + (F(1) - ax) * ay * t10 + ax * ay * t11).astype(F)


def _level(I, J, u, v, prevU, prevV, win, iters):
    """denseKernel for one level; u, v are updated in place where the reference writes."""
    h, w = I.shape
    wx_, wy_ = win
    hx, hy = (wx_ - 1) // 2, (wy_ - 1) // 2
    ys, xs = np.mgrid[0:h, 0:w]

    def T(dy, dx):
        return _texel(I, ys + dy, xs + dx)

    Ii = np.trunc(T(0, 0)).astype(np.int64)
    dIdx = np.trunc(3 * T(-1, 1) + 10 * T(0, 1) + 3 * T(1, 1) - (3 * T(-1, -1) + 10 * T(0, -1) + 3 * T(1, -1))).astype(np.int64)
    dIdy = np.trunc(3 * T(1, -1) + 10 * T(1, 0) + 3 * T(1, 1) - (3 * T(-1, -1) + 10 * T(-1, 0) + 3 * T(-1, 1))).astype(np.int64)

    def patch(a, i, j):   # value of patch array at window offset (i, j) for every pixel (clamped like the texture)
        return a[np.clip(ys - hy + i, 0, h - 1), np.clip(xs - hx + j, 0, w - 1)]

    # NB the patch entries for out-of-image positions are computed at clamped *texel* coordinates in
    # the reference (the Scharr stencil is evaluated around the clamped centre's neighbours, each
    # clamped individually); recompute them that way instead of clamping the derivative image.
    def patch_vals(i, j):
        py, px = ys - hy + i, xs - hx + j
        def TT(dy, dx):
            return _texel(I, py + dy, px + dx)
        Iv = np.trunc(TT(0, 0)).astype(np.int64)
        gx = np.trunc(3 * TT(-1, 1) + 10 * TT(0, 1) + 3 * TT(1, 1) - (3 * TT(-1, -1) + 10 * TT(0, -1) + 3 * TT(1, -1))).astype(np.int64)
        gy = np.trunc(3 * TT(1, -1) + 10 * TT(1, 0) + 3 * TT(1, 1) - (3 * TT(-1, -1) + 10 * TT(-1, 0) + 3 * TT(-1, 1))).astype(np.int64)
        return Iv, gx, gy

    pv = [[patch_vals(i, j) for j in range(wx_)] for i in range(wy_)]
    A11 = np.zeros((h, w), np.int64)
    A12 = np.zeros((h, w), np.int64)
    A22 = np.zeros((h, w), np.int64)
    for i in range(wy_):
        for j in range(wx_):
            _, gx, gy = pv[i][j]
            A11 = _wrap32(A11 + _wrap32(gx * gx))
            A12 = _wrap32(A12 + _wrap32(gx * gy))
            A22 = _wrap32(A22 + _wrap32(gy * gy))
    a11, a12, a22 = A11.astype(F), A12.astype(F), A22.astype(F)
    D = a11 * a22 - a12 * a12
    ok = ~(D < np.finfo(F).eps)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        Dinv = (F(1) / D).astype(F)
        a11, a12, a22 = (a11 * Dinv).astype(F), (a12 * Dinv).astype(F), (a22 * Dinv).astype(F)
    nx = (xs.astype(F) + prevU[ys // 2, xs // 2] * F(2)).astype(F)
    ny = (ys.astype(F) + prevV[ys // 2, xs // 2] * F(2)).astype(F)
    active = ok.copy()     # still iterating
    alive = ok.copy()      # will write
    for _ in range(iters):
        oob = (nx < 0) | (nx >= w) | (ny < 0) | (ny >= h)
        alive &= ~(active & oob)
        active &= ~oob
        if not active.any():
            break
        b1 = np.zeros((h, w), np.int64)
        b2 = np.zeros((h, w), np.int64)
        for i in range(wy_):
            jy = (ny - F(hy) + F(i) + F(0.5)).astype(F)
            for j in range(wx_):
                Iv, gx, gy = pv[i][j]
                Jv = np.trunc(_bilinear(J, jy, (nx - F(hx) + F(j) + F(0.5)).astype(F))).astype(np.int64)
                diff = _wrap32((Jv - Iv) * 32)
                b1 = _wrap32(b1 + _wrap32(diff * gx))
                b2 = _wrap32(b2 + _wrap32(diff * gy))
        fb1, fb2 = b1.astype(F), b2.astype(F)
        dx = (a12 * fb2 - a22 * fb1).astype(F)
        dy = (a12 * fb1 - a11 * fb2).astype(F)
        nx = np.where(active, nx + dx, nx).astype(F)
        ny = np.where(active, ny + dy, ny).astype(F)
        conv = (np.abs(dx) < F(0.01)) & (np.abs(dy) < F(0.01))
        active &= ~conv
    u[:h, :w] = np.where(alive, nx - xs.astype(F), u[:h, :w])
    v[:h, :w] = np.where(alive, ny - ys.astype(F), v[:h, :w])


def calc(I0: np.ndarray, I1: np.ndarray, winSize=(13, 13), maxLevel=3, iters=30) -> np.ndarray:
    """PyrLKOpticalFlowBase::dense + DensePyrLKOpticalFlowImpl::calc.  uint8 in, (H, W, 2) float32 out."""
    assert I0.dtype == np.uint8 and I0.shape == I1.shape
    h, w = I0.shape
    pI, pJ = [I0.astype(F)], [I1.astype(F)]
    for _ in range(maxLevel):
        pI.append(cv2.pyrDown(pI[-1]))   # same [1 4 6 4 1]/16 kernel and REFLECT101 border as cuda::pyrDown
        pJ.append(cv2.pyrDown(pJ[-1]))
    uP = [np.zeros((h, w), F), np.zeros((h, w), F)]
    vP = [np.zeros((h, w), F), np.zeros((h, w), F)]
    idx = 0
    for level in range(maxLevel, -1, -1):
        idx2 = (idx + 1) & 1
        _level(pI[level], pJ[level], uP[idx], vP[idx], uP[idx2], vP[idx2], winSize, iters)
        if level > 0:
            idx = idx2
    return np.stack([uP[idx], vP[idx]], axis=-1)
