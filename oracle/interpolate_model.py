"""numpy model of ``cv::cuda::interpolateFrames`` (modules/cudalegacy/src/interpolate_frames.cpp:54-111,
src/cuda/NPP_staging.cu:1648-1790 nppiStInterpolateFrames / BlendFramesKernel, :1838-1905
ForwardWarpKernel_PSF2x2, :1956-1996 NormalizeKernel / MemsetKernel, :2022-2063 nppiStVectorWarp_PSF2x2).

TEST INFRASTRUCTURE (see oracle/__init__.py).  **Parity unpinned**: the reference has no CPU
implementation and its only test (cudalegacy test_nvidia / NCV) needs opencv_extra data.  This model
follows the reference call by call -- four sequential splat passes with their coverage clears, then the
blend -- and therefore reproduces its defects when ``corrected=False``:
  * the 4th pass writes bwdU again (NPP_staging.cu:1779-1787), bwdV stays zero;
  * the "visible in both frames" branch samples frame0 twice (:1666);
  * MemsetKernel clears ``i*w + j`` (:1985-1996), i.e. only the first w*h floats of a pitched plane.
Float atomics make the GPU sum order arbitrary; the model accumulates in raster order, so comparisons
use a rounding-level tolerance.  The hardware bilinear filter is modelled with 8-bit fractional weights.
"""
from __future__ import annotations

import numpy as np

F = np.float32


def _splat(src, u, v, time_scale, norm, dst, w, h, s):
    """ForwardWarpKernel_PSF2x2: scatter ``src`` into flat pitched ``dst`` / ``norm`` (length s*h)."""
    jj, ii = np.meshgrid(np.arange(w), np.arange(h))
    cx = (u * F(time_scale) + jj.astype(F) + F(1)).astype(F)
    cy = (v * F(time_scale) + ii.astype(F) + F(1)).astype(F)
    px, py = np.trunc(cx), np.trunc(cy)  # modff: integral part truncates toward zero
    dx, dy = (cx - px).astype(F), (cy - py).astype(F)
    tx0, ty0 = px.astype(np.int64), py.astype(np.int64)
    one = F(1)
    taps = [(tx0, ty0, dx * dy), (tx0 - 1, ty0, (one - dx) * dy), (tx0 - 1, ty0 - 1, (one - dx) * (one - dy)),
            (tx0, ty0 - 1, dx * (one - dy))]
    for tx, ty, wgt in taps:
        ok = (tx >= 0) & (tx < w) & (ty >= 0) & (ty < h)
        o = (ty * s + tx)[ok]
        wv = wgt[ok].astype(F)
        # float32 accumulation in raster order, like a serialised stream of atomics
        _add_at_f32(dst, o, (src[ok] * wv).astype(F))
        _add_at_f32(norm, o, wv)


def _add_at_f32(dst, idx, val):
    # np.add.at accumulates in the array dtype (float32), element by element
    np.add.at(dst, idx, val)


def _vector_warp(src, u, v, time_scale, norm, dst, w, h, s):
    """nppiStVectorWarp_PSF2x2_32f_C1: clear (by width!), splat, normalise."""
    norm[: w * h] = 0  # MemsetKernel indexes i*w + j
    _splat(src, u, v, time_scale, norm, dst, w, h, s)
    inv = np.where(norm == 0, F(1), (F(1) / np.where(norm == 0, F(1), norm))).astype(F)
    # NormalizeKernel touches columns < w only
    d2, i2 = dst.reshape(h, s), inv.reshape(h, s)
    d2[:, :w] = (d2[:, :w] * i2[:, :w]).astype(F)


def _tex_linear(img, y, x):
    h, w = img.shape
    xb, yb = (x - F(0.5)).astype(F), (y - F(0.5)).astype(F)
    fx, fy = np.floor(xb), np.floor(yb)
    ax = (np.floor((xb - fx) * F(256) + F(0.5)) * F(1 / 256)).astype(F)
    ay = (np.floor((yb - fy) * F(256) + F(0.5)) * F(1 / 256)).astype(F)
    x0 = np.clip(fx, -2, w + 1).astype(np.int64)
    y0 = np.clip(fy, -2, h + 1).astype(np.int64)
    xa, xc = np.clip(x0, 0, w - 1), np.clip(x0 + 1, 0, w - 1)
    ya, yc = np.clip(y0, 0, h - 1), np.clip(y0 + 1, 0, h - 1)
    one = F(1)
    return ((one - ax) * (one - ay) * img[ya, xa] + ax * (one - ay) * img[ya, xc] + (one - ax) * ay * img[yc, xa] +
            ax * ay * img[yc, xc]).astype(F)


def interpolate_frames(frame0, frame1, fu, fv, bu, bv, pos, stride=None, corrected=False):
    """Returns (new_frame (h, w), buf (6, h, stride)) with buf = cov0, cov1, fwdU, fwdV, bwdU, bwdV."""
    h, w = frame0.shape
    s = w if stride is None else int(stride)
    pos = F(pos)
    f = [np.ascontiguousarray(a, F) for a in (frame0, frame1, fu, fv, bu, bv)]
    frame0, frame1, fu, fv, bu, bv = f
    buf = np.zeros((6, h * s), F)  # buf.setTo(0)
    cov0, cov1, fwdU, fwdV, bwdU, bwdV = buf
    if corrected:
        cov_a, cov_b = np.zeros(h * s, F), np.zeros(h * s, F)
        _splat(fu, fu, fv, pos, cov0, fwdU, w, h, s)
        _splat(fv, fu, fv, pos, cov_a, fwdV, w, h, s)
        _splat(bu, bu, bv, F(1) - pos, cov1, bwdU, w, h, s)
        _splat(bv, bu, bv, F(1) - pos, cov_b, bwdV, w, h, s)
        for dst, cov in ((fwdU, cov0), (fwdV, cov0), (bwdU, cov1), (bwdV, cov1)):
            inv = np.where(cov == 0, F(1), F(1) / np.where(cov == 0, F(1), cov)).astype(F)
            dst *= inv
    else:
        _vector_warp(fu, fu, fv, pos, cov0, fwdU, w, h, s)
        _vector_warp(fv, fu, fv, pos, cov0, fwdV, w, h, s)
        _vector_warp(bu, bu, bv, F(1) - pos, cov1, bwdU, w, h, s)
        _vector_warp(bv, bu, bv, F(1) - pos, cov1, bwdU, w, h, s)  # sic: bwdU again

    B = buf.reshape(6, h, s)
    u, v, ur, vr = B[2, :, :w], B[3, :, :w], B[4, :, :w], B[5, :, :w]
    o0, o1 = B[0, :, :w], B[1, :, :w]
    jj, ii = np.meshgrid(np.arange(w), np.arange(h))
    x, y = jj.astype(F) + F(0.5), ii.astype(F) + F(0.5)
    one = F(1)
    b0, b1 = o0 > F(1e-4), o1 > F(1e-4)
    a = _tex_linear(frame0, y - v * pos, x - u * pos)
    second = frame1 if corrected else frame0
    bsample = _tex_linear(second, y + v * (one - pos), x + u * (one - pos))
    both = (a * (one - pos) + bsample * pos).astype(F)
    c = _tex_linear(frame1, y - vr * (one - pos), x - ur * (one - pos))
    out = np.where(b0 & b1, both, np.where(b0, a, c)).astype(F)
    return out, B
