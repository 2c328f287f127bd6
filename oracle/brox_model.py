"""numpy model of the reference's Brox et al. optical flow, ``cv::cuda::BroxOpticalFlow``
(modules/cudaoptflow/src/brox.cpp -> modules/cudalegacy/src/cuda/NCVBroxOpticalFlow.cu +
NPP_staging.cu filters / resizers).

TEST INFRASTRUCTURE (see oracle/__init__.py).  **Parity unpinned**: the reference has no CPU
implementation of Brox, its golden vector (opencv_extra opticalflow/brox_optical_flow_cc20.bin) is
not in the container and the CUDA reference cannot be built here; this restatement is checked only
for self-consistency (recovers synthetic ground truth, no NaNs) and the CUDA engine is checked
against it.  One documented deviation: the reference samples images through hardware bilinear
texture filtering (9-bit fixed-point weights); the model (and the engine) use exact float32 weights.

Citations: NB = /root/reference/modules/cudalegacy/src/cuda/NCVBroxOpticalFlow.cu,
           NS = /root/reference/modules/cudalegacy/src/cuda/NPP_staging.cu.
"""
from __future__ import annotations

import math

import numpy as np

F = np.float32
EPS2 = F(1e-6)  # NB:78


class BroxParams:
    """Defaults of cv::cuda::BroxOpticalFlow::create (cudaoptflow.hpp:179-185)."""

    def __init__(self, alpha=0.197, gamma=50.0, scale_factor=0.8, inner_iterations=5, outer_iterations=150,
                 solver_iterations=10):
        self.alpha, self.gamma, self.scale_factor = alpha, gamma, scale_factor
        self.inner_iterations, self.outer_iterations, self.solver_iterations = (
            inner_iterations, outer_iterations, solver_iterations)


def pyramid_sizes(h: int, w: int, scale_factor: float, outer_iterations: int):
    """NB:730-785: ceilf(src * scale), cumulative float32 scale, stop at <= 15 px or the level cap."""
    sizes = [(h, w)]
    sf = F(scale_factor)
    scale = F(1.0) * sf
    pw, ph = w, h
    while pw > 15 and ph > 15 and len(sizes) < outer_iterations:
        lw = int(math.ceil(float(F(w) * scale)))
        lh = int(math.ceil(float(F(h) * scale)))
        sizes.append((lh, lw))
        scale = F(scale * sf)
        pw, ph = lw, lh
    return sizes


def resize_supersample(src: np.ndarray, dh: int, dw: int, scale: float) -> np.ndarray:
    """resizeSuperSample_32f (NS:2073-2149) with scaleX = scaleY = 1/factor."""
    sh, sw = src.shape
    s = F(scale)
    out = np.zeros((dh, dw), F)

    def ranges(n_dst, n_src):
        x = s * np.arange(n_dst, dtype=F)
        b = np.maximum(x - s, F(0)).astype(F)
        e = np.minimum(x + s, F(n_src - 1)).astype(F)
        fb = np.floor(b).astype(F)
        ce = np.ceil(e).astype(F)
        return b, e, fb, ce, fb.astype(np.int64), ce.astype(np.int64)

    xb, xe, fxb, cxe, ixb, ixe = ranges(dw, sw)
    yb, ye, fyb, cye, iyb, iye = ranges(dh, sh)

    def line(row, j):
        # processLine for every dst x of src row `row`
        wsum = F(1.0) - xb + fxb
        acc = row[ixb] * (F(1.0) - xb + fxb)
        kmax = int((ixe - ixb).max())
        for k in range(1, kmax + 1):
            pos = ixb + k
            inner = pos < ixe
            last = pos == ixe
            v = row[np.minimum(pos, sw - 1)]
            acc = np.where(inner, acc + v, acc)
            wsum = np.where(inner, wsum + F(1.0), wsum)
            acc = np.where(last, acc + v * (cxe - xe), acc)
            wsum = np.where(last, wsum + (cxe - xe), wsum)
        # ixe == ixb: the loop body never runs but the "last" term is still added at spos = ixb + 1
        deg = ixe == ixb
        if deg.any():
            v = row[np.minimum(ixb + 1, sw - 1)]
            acc = np.where(deg, acc + v * (cxe - xe), acc)
            wsum = np.where(deg, wsum + (cxe - xe), wsum)
        return (acc / wsum).astype(F)

    for y in range(dh):
        wsum = F(1.0) - yb[y] + fyb[y]
        acc = line(src[iyb[y]], y) * (F(1.0) - yb[y] + fyb[y])
        yy = iyb[y] + 1
        while yy < iye[y]:
            acc = acc + line(src[yy], y)
            wsum = wsum + F(1.0)
            yy += 1
        acc = acc + line(src[min(yy, sh - 1)], y) * (cye[y] - ye[y])
        wsum = wsum + (cye[y] - ye[y])
        out[y] = acc / wsum
    return out


def _mirror_filter_idx(i, n):
    """getValueMirrorRow/Column (NS:1433-1447): i < 0 -> 1 - i (asymmetric!), i >= n -> 2n - i - 1."""
    i = np.where(i < 0, 1 - i, i)
    i = np.where(i >= n, 2 * n - i - 1, i)
    return np.clip(i, 0, n - 1)


def filter5(src: np.ndarray, axis: int) -> np.ndarray:
    """FilterRow/ColumnBorderMirror_32f_C1R with {1,-8,0,8,-1} * 1/12, anchor 2 (NS:1449-1501; NB:843-868)."""
    k = np.array([1.0, -8.0, 0.0, 8.0, -1.0], F)
    n = src.shape[axis]
    idx = np.arange(n)
    acc = np.zeros_like(src)
    for m in range(5):
        j = _mirror_filter_idx(idx + m - 2, n)
        acc = acc + np.take(src, j, axis=axis) * k[m]
    return (acc * F(1.0 / 12.0)).astype(F)


def _mirror_tex(i, n):
    """cudaAddressModeMirror on texel indices: -1 -> 0, -2 -> 1, n -> n-1, n+1 -> n-2 (period 2n)."""
    i = np.mod(i, 2 * n)
    return np.where(i >= n, 2 * n - 1 - i, i)


def tex_bilinear_mirror(img: np.ndarray, xn, yn):
    """Normalised-coordinate, linear-filtered, mirror-addressed texture fetch (NB:829-830,870-876),
    with exact float32 weights."""
    h, w = img.shape
    xb = xn * F(w) - F(0.5)
    yb = yn * F(h) - F(0.5)
    x0 = np.floor(xb)
    y0 = np.floor(yb)
    ax = (xb - x0).astype(F)
    ay = (yb - y0).astype(F)
    x0 = x0.astype(np.int64)
    y0 = y0.astype(np.int64)
    xa, xc = _mirror_tex(x0, w), _mirror_tex(x0 + 1, w)
    ya, yc = _mirror_tex(y0, h), _mirror_tex(y0 + 1, h)
    top = img[ya, xa] * (F(1) - ax) + img[ya, xc] * ax
    bot = img[yc, xa] * (F(1) - ax) + img[yc, xc] * ax
    return (top * (F(1) - ay) + bot * ay).astype(F)


def resize_bicubic(src: np.ndarray, dh: int, dw: int, scale: float) -> np.ndarray:
    """resizeBicubic (NS:2172-2231): taps ceil(x-2)..floor(x+2) clamped to the ROI, Keys a=-0.5,
    sum / wsum, point texture with mirror addressing (never reached because of the clamp)."""
    from .tvl1_gpu_model import bicubic_coeff
    sh, sw = src.shape
    s = F(scale)
    ys, xs = np.mgrid[0:dh, 0:dw]
    x = s * xs.astype(F)
    y = s * ys.astype(F)
    xmin = np.maximum(np.ceil(x - F(2)), F(0))
    xmax = np.minimum(np.floor(x + F(2)), F(sw - 1))
    ymin = np.maximum(np.ceil(y - F(2)), F(0))
    ymax = np.minimum(np.floor(y + F(2)), F(sh - 1))
    acc = np.zeros((dh, dw), F)
    wsum = np.zeros((dh, dw), F)
    for j in range(5):
        cy = ymin + F(j)
        my = cy <= ymax
        wy = bicubic_coeff(y - cy)
        cyi = np.clip(cy.astype(np.int64), 0, sh - 1)
        for i in range(5):
            cx = xmin + F(i)
            m = my & (cx <= xmax)
            wgt = np.where(m, bicubic_coeff(x - cx) * wy, F(0)).astype(F)
            cxi = np.clip(cx.astype(np.int64), 0, sw - 1)
            acc = acc + wgt * src[cyi, cxi]
            wsum = wsum + wgt
    with np.errstate(divide="ignore", invalid="ignore"):
        out = np.where(wsum == 0, F(0), acc / wsum)
    return out.astype(F)


def _mirror_load(a: np.ndarray, dy: int, dx: int) -> np.ndarray:
    """load_array_element (NB:238-249): neighbour (i+dx, j+dy) with mirror i<0 -> -i-1, i>=w -> 2w-i-1."""
    h, w = a.shape
    ii = np.arange(w) + dx
    jj = np.arange(h) + dy
    ii = np.maximum(ii, -ii - 1)
    ii = np.minimum(ii, w - ii + w - 1)
    jj = np.maximum(jj, -jj - 1)
    jj = np.minimum(jj, h - jj + h - 1)
    return a[np.ix_(jj, ii)]


def prepare_sor(u, v, du, dv, I0, I1, Ix, Ixx, Ix0, Iy, Iyy, Iy0, Ixy, alpha, gamma):
    """prepare_sor_stage_1_tex + prepare_sor_stage_2 (NB:340-473)."""
    h, w = u.shape
    alpha, gamma = F(alpha), F(gamma)
    jg, ig = np.mgrid[0:h, 0:w]
    x = ig.astype(F) + F(0.5)
    y = jg.astype(F) + F(0.5)
    wx = (x + u) / F(w)
    wy = (y + v) / F(h)
    xn = x / F(w)
    yn = y / F(h)
    Iz = tex_bilinear_mirror(I1, wx, wy) - tex_bilinear_mirror(I0, xn, yn)
    ix = tex_bilinear_mirror(Ix, wx, wy)
    ixz = ix - tex_bilinear_mirror(Ix0, xn, yn)
    ixy = tex_bilinear_mirror(Ixy, wx, wy)
    ixx = tex_bilinear_mirror(Ixx, wx, wy)
    iy = tex_bilinear_mirror(Iy, wx, wy)
    iyz = iy - tex_bilinear_mirror(Iy0, xn, yn)
    iyy = tex_bilinear_mirror(Iyy, wx, wy)
    q0 = Iz + ix * du + iy * dv
    q1 = ixz + ixx * du + ixy * dv
    q2 = iyz + ixy * du + iyy * dv
    data = (F(0.5) / np.sqrt(q0 * q0 + gamma * (q1 * q1 + q2 * q2) + EPS2)).astype(F)
    data = (data / alpha).astype(F)

    U, V = (u + du).astype(F), (v + dv).astype(F)  # u[pos] + du[pos] ... (NB:192-193) -- summed pairwise below
    # diffusivity_along_x (NB:188-204): derivative between (i,j) and (i-1,j)
    def n(a, dy, dx):
        return _mirror_load(a, dy, dx)
    u_x = u + du - n(u, 0, -1) - n(du, 0, -1)
    v_x = v + dv - n(v, 0, -1) - n(dv, 0, -1)
    u_y = F(0.25) * (n(u, 1, 0) + n(du, 1, 0) + n(u, 1, -1) + n(du, 1, -1) - n(u, -1, 0) - n(du, -1, 0)
                     - n(u, -1, -1) - n(du, -1, -1))
    v_y = F(0.25) * (n(v, 1, 0) + n(dv, 1, 0) + n(v, 1, -1) + n(dv, 1, -1) - n(v, -1, 0) - n(dv, -1, 0)
                     - n(v, -1, -1) - n(dv, -1, -1))
    sx = (F(0.5) / np.sqrt(u_x * u_x + v_x * v_x + u_y * u_y + v_y * v_y + EPS2)).astype(F)
    # diffusivity_along_y (NB:216-227): derivative between (i,j) and (i,j-1)
    u_y = u + du - n(u, -1, 0) - n(du, -1, 0)
    v_y = v + dv - n(v, -1, 0) - n(dv, -1, 0)
    u_x = F(0.25) * (n(u, 0, 1) + n(u, -1, 1) + n(du, 0, 1) + n(du, -1, 1) - n(u, 0, -1) - n(u, -1, -1)
                     - n(du, 0, -1) - n(du, -1, -1))
    v_x = F(0.25) * (n(v, 0, 1) + n(v, -1, 1) + n(dv, 0, 1) + n(dv, -1, 1) - n(v, 0, -1) - n(v, -1, -1)
                     - n(dv, 0, -1) - n(dv, -1, -1))
    sy = (F(0.5) / np.sqrt(u_x * u_x + v_x * v_x + u_y * u_y + v_y * v_y + EPS2)).astype(F)
    sx[:, 0] = 0
    sy[0, :] = 0
    num_dudv = data * (ix * iy + gamma * ixy * (ixx + iyy))
    num_u = data * (ix * Iz + gamma * (ixx * ixz + ixy * iyz))
    num_v = data * (iy * Iz + gamma * (iyy * iyz + ixy * ixz))
    den_u = data * (ix * ix + gamma * (ixy * ixy + ixx * ixx))
    den_v = data * (iy * iy + gamma * (ixy * ixy + iyy * iyy))
    # stage 2: + sx(i) + sx(i+1) + sy(j) + sy(j+1), zero beyond the image
    sxr = np.zeros_like(sx)
    sxr[:, :-1] = sx[:, 1:]
    syu = np.zeros_like(sy)
    syu[:-1, :] = sy[1:, :]
    ssum = sx + sxr + sy + syu
    inv_u = (F(1.0) / (den_u + ssum)).astype(F)
    inv_v = (F(1.0) / (den_v + ssum)).astype(F)
    return sx, sy, inv_u, inv_v, num_dudv.astype(F), num_u.astype(F), num_v.astype(F)


def sor_pass(is_black, u, v, du, dv, sx, sy, inv_u, inv_v, num_u, num_v, num_dudv, omega=F(1.99)):
    """sor_pass<isBlack> (NB:479-554): neighbours clamped at borders, s_right / s_up zero at the far edges."""
    h, w = u.shape
    jj, ii = np.mgrid[0:h, 0:w]

    def sh(a, dy, dx):
        return a[np.clip(jj + dy, 0, h - 1), np.clip(ii + dx, 0, w - 1)]

    s_left, s_down = sx, sy
    s_right = np.where(ii < w - 1, sh(sx, 0, 1), F(0)).astype(F)
    s_up = np.where(jj < h - 1, sh(sy, 1, 0), F(0)).astype(F)
    ssum = s_left + s_right + s_up + s_down
    numer_u = (s_left * (sh(u, 0, -1) + sh(du, 0, -1)) + s_up * (sh(u, 1, 0) + sh(du, 1, 0))
               + s_right * (sh(u, 0, 1) + sh(du, 0, 1)) + s_down * (sh(u, -1, 0) + sh(du, -1, 0))
               - u * ssum - num_u - num_dudv * dv)
    du_new = (F(1.0) - omega) * du + omega * inv_u * numer_u
    numer_v = (s_left * (sh(v, 0, -1) + sh(dv, 0, -1)) + s_up * (sh(v, 1, 0) + sh(dv, 1, 0))
               + s_right * (sh(v, 0, 1) + sh(dv, 0, 1)) + s_down * (sh(v, -1, 0) + sh(dv, -1, 0))
               - v * ssum - num_v - num_dudv * du_new)
    dv_new = (F(1.0) - omega) * dv + omega * inv_v * numer_v
    m = ((ii + jj) % 2) == is_black
    return np.where(m, du_new, du).astype(F), np.where(m, dv_new, dv).astype(F)


def calc(I0: np.ndarray, I1: np.ndarray, P: BroxParams | None = None) -> np.ndarray:
    """NCVBroxOpticalFlow (NB:598-985) behind BroxOpticalFlowImpl::calc (brox.cpp:129-188).
    I0, I1: float32 in [0, 1].  Returns flow (H, W, 2)."""
    P = P or BroxParams()
    assert I0.dtype == np.float32 and I0.shape == I1.shape
    h, w = I0.shape
    sizes = pyramid_sizes(h, w, P.scale_factor, P.outer_iterations)
    pyr0, pyr1 = [I0], [I1]
    inv_sf = F(1.0) / F(P.scale_factor)
    for (lh, lw) in sizes[1:]:
        pyr0.append(resize_supersample(pyr0[-1], lh, lw, inv_sf))
        pyr1.append(resize_supersample(pyr1[-1], lh, lw, inv_sf))
    lh, lw = sizes[-1]
    u = np.zeros((lh, lw), F)
    v = np.zeros((lh, lw), F)
    for lvl in range(len(sizes) - 1, -1, -1):
        A, B = pyr0[lvl], pyr1[lvl]
        du = np.zeros_like(u)
        dv = np.zeros_like(v)
        Ix0, Iy0 = filter5(A, 1), filter5(A, 0)
        Ix, Iy = filter5(B, 1), filter5(B, 0)
        Ixx, Iyy, Ixy = filter5(Ix, 1), filter5(Iy, 0), filter5(Iy, 1)
        for _ in range(P.inner_iterations):
            sx, sy, inv_u, inv_v, num_dudv, num_u, num_v = prepare_sor(u, v, du, dv, A, B, Ix, Ixx, Ix0, Iy, Iyy,
                                                                         Iy0, Ixy, P.alpha, P.gamma)
            for _ in range(P.solver_iterations):
                du, dv = sor_pass(0, u, v, du, dv, sx, sy, inv_u, inv_v, num_u, num_v, num_dudv)
                du, dv = sor_pass(1, u, v, du, dv, sx, sy, inv_u, inv_v, num_u, num_v, num_dudv)
        u = (u + du).astype(F)
        v = (v + dv).astype(F)
        if lvl > 0:
            nh, nw = sizes[lvl - 1]
            s = F(1.0) / inv_sf  # kernel gets 1/xFactor with xFactor = 1/scale_factor  (NB:952-953, NS:2270)
            u = (resize_bicubic(u, nh, nw, s) * inv_sf).astype(F)
            v = (resize_bicubic(v, nh, nw, s) * inv_sf).astype(F)
    return np.stack([u, v], axis=-1)
