"""CPU oracle for Dual TV-L1: a restatement of the reference's CPU path
``cv::optflow::DualTVL1OpticalFlow`` (modules/optflow/src/tvl1flow.cpp).

TEST INFRASTRUCTURE (see oracle/__init__.py) -- never imported by the product.

Pinning status: the opencv_contrib python module (cv2.optflow) is not in this
image and the reference C++ cannot be compiled here (no OpenCV core headers),
and the reference's golden vector (opencv_extra optflow/tvl1_flow.flo) is not
in the container.  The restatement below calls the *reference's own external
primitives* through cv2 4.13 -- cv::resize(INTER_LINEAR), cv::remap(INTER_CUBIC),
cv::medianBlur -- exactly where tvl1flow.cpp calls them, and restates only the
in-tree arithmetic.  End-to-end parity is therefore "pinned on primitives,
unpinned end to end"; DESIGN.md says so.

Each function cites the reference lines it follows (paths relative to
/root/reference/modules/optflow/src/tvl1flow.cpp).
"""
from __future__ import annotations

import numpy as np
import cv2

F = np.float32
FLT_EPS = np.finfo(np.float32).eps
FLT_MAX = np.finfo(np.float32).max


class TVL1Params:
    """Defaults of OpticalFlowDual_TVL1::OpticalFlowDual_TVL1() (:387-400)."""

    def __init__(self, tau=0.25, lambda_=0.15, theta=0.3, nscales=5, warps=5, epsilon=0.01,
                 innerIterations=30, outerIterations=10, scaleStep=0.8, gamma=0.0,
                 medianFiltering=5, useInitialFlow=False):
        self.tau, self.lambda_, self.theta = tau, lambda_, theta
        self.nscales, self.warps, self.epsilon = nscales, warps, epsilon
        self.innerIterations, self.outerIterations = innerIterations, outerIterations
        self.scaleStep, self.gamma = scaleStep, gamma
        self.medianFiltering, self.useInitialFlow = medianFiltering, useInitialFlow


def centered_gradient(src):
    """:718-770 (body :697-716). 0.5*central difference; one-sided *0.5* at borders."""
    dx = np.empty_like(src)
    dy = np.empty_like(src)
    dx[:, 1:-1] = F(0.5) * (src[:, 2:] - src[:, :-2])
    dx[:, 0] = F(0.5) * (src[:, 1] - src[:, 0])
    dx[:, -1] = F(0.5) * (src[:, -1] - src[:, -2])
    dy[1:-1, :] = F(0.5) * (src[2:, :] - src[:-2, :])
    dy[0, :] = F(0.5) * (src[1, :] - src[0, :])
    dy[-1, :] = F(0.5) * (src[-1, :] - src[-2, :])
    return dx, dy


def forward_gradient(src):
    """:804-840. Forward differences, zero on the last row/column."""
    dx = np.zeros_like(src)
    dy = np.zeros_like(src)
    dx[:, :-1] = src[:, 1:] - src[:, :-1]
    dy[:-1, :] = src[1:, :] - src[:-1, :]
    return dx, dy


def divergence(v1, v2):
    """:874-899 (body :855-872). Backward differences, first row/col special-cased."""
    div = np.empty_like(v1)
    div[1:, 1:] = (v1[1:, 1:] - v1[1:, :-1]) + (v2[1:, 1:] - v2[:-1, 1:])
    div[0, 1:] = v1[0, 1:] - v1[0, :-1] + v2[0, 1:]
    div[1:, 0] = v1[1:, 0] + v2[1:, 0] - v2[:-1, 0]
    div[0, 0] = v1[0, 0] + v2[0, 0]
    return div


def calc_grad_rho(I0, I1w, I1wx, I1wy, u1, u2):
    """:920-944."""
    grad = I1wx * I1wx + I1wy * I1wy
    rho_c = I1w - I1wx * u1 - I1wy * u2 - I0
    return grad, rho_c


def estimate_v(I1wx, I1wy, u1, u2, u3, grad, rho_c, l_t, gamma):
    """:992-1041 thresholding operator TH."""
    use_gamma = gamma != 0
    l_t = F(l_t)
    gamma = F(gamma)
    rho = rho_c + (I1wx * u1 + I1wy * u2)
    if use_gamma:
        rho = rho + gamma * u3
    c1 = rho < -l_t * grad
    c2 = (~c1) & (rho > l_t * grad)
    c3 = (~c1) & (~c2) & (grad > FLT_EPS)
    with np.errstate(divide="ignore", invalid="ignore"):
        fi = np.where(c3, -rho / np.where(c3, grad, F(1)), F(0)).astype(F)
    d1 = np.where(c1, l_t * I1wx, np.where(c2, -l_t * I1wx, np.where(c3, fi * I1wx, F(0)))).astype(F)
    d2 = np.where(c1, l_t * I1wy, np.where(c2, -l_t * I1wy, np.where(c3, fi * I1wy, F(0)))).astype(F)
    v1 = u1 + d1
    v2 = u2 + d2
    v3 = None
    if use_gamma:
        d3 = np.where(c1, l_t * gamma, np.where(c2, -l_t * gamma, np.where(c3, fi * gamma, F(0)))).astype(F)
        v3 = u3 + d3
    return v1, v2, v3


def estimate_u(v1, v2, v3, div_p1, div_p2, div_p3, u1, u2, u3, theta, gamma):
    """:1074-1116.  Returns (u1, u2, u3, error); error is the reference's *serial float32*
    accumulation in row-major order (np.cumsum reproduces that order)."""
    theta = F(theta)
    n1 = v1 + theta * div_p1
    n2 = v2 + theta * div_p2
    e = (n1 - u1) * (n1 - u1) + (n2 - u2) * (n2 - u2)
    n3 = None
    if gamma != 0:
        n3 = v3 + theta * div_p3
        e = e + (n3 - u3) * (n3 - u3)
    err = float(np.cumsum(e.ravel(), dtype=F)[-1])
    return n1, n2, n3, err


def estimate_dual(u1x, u1y, u2x, u2y, u3x, u3y, p11, p12, p21, p22, p31, p32, taut, use_gamma):
    """:1140-1180.  hypot in double then cast (static_cast<float>(hypot(..)))."""
    taut = F(taut)
    g1 = np.hypot(u1x.astype(np.float64), u1y.astype(np.float64)).astype(F)
    g2 = np.hypot(u2x.astype(np.float64), u2y.astype(np.float64)).astype(F)
    ng1 = F(1.0) + taut * g1
    ng2 = F(1.0) + taut * g2
    p11 = (p11 + taut * u1x) / ng1
    p12 = (p12 + taut * u1y) / ng1
    p21 = (p21 + taut * u2x) / ng2
    p22 = (p22 + taut * u2y) / ng2
    if use_gamma:
        g3 = np.hypot(u3x.astype(np.float64), u3y.astype(np.float64)).astype(F)
        ng3 = F(1.0) + taut * g3
        p31 = (p31 + taut * u3x) / ng3
        p32 = (p32 + taut * u3y) / ng3
    return p11, p12, p21, p22, p31, p32


def proc_one_scale(P: TVL1Params, I0, I1, u1, u2, u3):
    """:1313-1408."""
    h, w = I0.shape
    scaledEpsilon = float(F(P.epsilon * P.epsilon * (h * w)))
    use_gamma = P.gamma != 0.0
    I1x, I1y = centered_gradient(I1)
    z = lambda: np.zeros((h, w), F)
    p11, p12, p21, p22 = z(), z(), z(), z()
    p31 = p32 = None
    if use_gamma:
        p31, p32 = z(), z()
    l_t = F(P.lambda_ * P.theta)
    taut = F(P.tau / P.theta)
    ys, xs = np.mgrid[0:h, 0:w].astype(F)
    for _ in range(P.warps):
        map1 = xs + u1  # buildFlowMap :651-668
        map2 = ys + u2
        I1w = cv2.remap(I1, map1, map2, cv2.INTER_CUBIC)  # :1371-1373
        I1wx = cv2.remap(I1x, map1, map2, cv2.INTER_CUBIC)
        I1wy = cv2.remap(I1y, map1, map2, cv2.INTER_CUBIC)
        grad, rho_c = calc_grad_rho(I0, I1w, I1wx, I1wy, u1, u2)
        error = float(FLT_MAX)
        n_outer = 0
        while error > scaledEpsilon and n_outer < P.outerIterations:
            if P.medianFiltering > 1:
                u1 = cv2.medianBlur(u1, P.medianFiltering)
                u2 = cv2.medianBlur(u2, P.medianFiltering)
            n_inner = 0
            while error > scaledEpsilon and n_inner < P.innerIterations:
                v1, v2, v3 = estimate_v(I1wx, I1wy, u1, u2, u3, grad, rho_c, l_t, P.gamma)
                div_p1 = divergence(p11, p12)
                div_p2 = divergence(p21, p22)
                div_p3 = divergence(p31, p32) if use_gamma else None
                u1, u2, u3n, error = estimate_u(v1, v2, v3, div_p1, div_p2, div_p3, u1, u2, u3,
                                                P.theta, P.gamma)
                if use_gamma:
                    u3 = u3n
                u1x, u1y = forward_gradient(u1)
                u2x, u2y = forward_gradient(u2)
                u3x = u3y = None
                if use_gamma:
                    u3x, u3y = forward_gradient(u3)
                p11, p12, p21, p22, p31, p32 = estimate_dual(
                    u1x, u1y, u2x, u2y, u3x, u3y, p11, p12, p21, p22, p31, p32, taut, use_gamma)
                n_inner += 1
            n_outer += 1
    return u1, u2, u3


def calc(I0: np.ndarray, I1: np.ndarray, P: TVL1Params | None = None, init_flow=None) -> np.ndarray:
    """OpticalFlowDual_TVL1::calc (:402-533).  I0/I1: uint8 (kept 0..255) or float32 (x255).
    Returns flow (H, W, 2) float32."""
    P = P or TVL1Params()
    assert I0.dtype in (np.uint8, np.float32) and I0.shape == I1.shape and I0.dtype == I1.dtype
    assert P.nscales > 0
    k = F(1.0) if I0.dtype == np.uint8 else F(255.0)
    I0s = [I0.astype(F) * k]
    I1s = [I1.astype(F) * k]
    use_gamma = P.gamma != 0.0
    h, w = I0.shape
    u1s = [np.zeros((h, w), F)]
    u2s = [np.zeros((h, w), F)]
    if P.useInitialFlow:
        assert init_flow is not None and init_flow.shape == (h, w, 2)
        u1s[0] = init_flow[..., 0].astype(F).copy()
        u2s[0] = init_flow[..., 1].astype(F).copy()
    nscales = P.nscales
    for s in range(1, nscales):
        a = cv2.resize(I0s[s - 1], None, fx=P.scaleStep, fy=P.scaleStep, interpolation=cv2.INTER_LINEAR)
        b = cv2.resize(I1s[s - 1], None, fx=P.scaleStep, fy=P.scaleStep, interpolation=cv2.INTER_LINEAR)
        I0s.append(a)
        I1s.append(b)
        if a.shape[1] < 16 or a.shape[0] < 16:
            nscales = s
            break
        if P.useInitialFlow:
            u1s.append(cv2.resize(u1s[s - 1], None, fx=P.scaleStep, fy=P.scaleStep,
                                  interpolation=cv2.INTER_LINEAR) * F(P.scaleStep))
            u2s.append(cv2.resize(u2s[s - 1], None, fx=P.scaleStep, fy=P.scaleStep,
                                  interpolation=cv2.INTER_LINEAR) * F(P.scaleStep))
        else:
            u1s.append(np.zeros(a.shape, F))
            u2s.append(np.zeros(a.shape, F))
    u3 = np.zeros(I0s[nscales - 1].shape, F) if use_gamma else None
    for s in range(nscales - 1, -1, -1):
        u1s[s], u2s[s], u3 = proc_one_scale(P, I0s[s], I1s[s], u1s[s], u2s[s], u3)
        if s == 0:
            break
        hh, ww = I0s[s - 1].shape
        inv = F(1.0 / P.scaleStep)
        u1s[s - 1] = cv2.resize(u1s[s], (ww, hh), interpolation=cv2.INTER_LINEAR) * inv
        u2s[s - 1] = cv2.resize(u2s[s], (ww, hh), interpolation=cv2.INTER_LINEAR) * inv
        if use_gamma:
            u3 = cv2.resize(u3, (ww, hh), interpolation=cv2.INTER_LINEAR)  # not rescaled (:526-529)
    return np.stack([u1s[0], u2s[0]], axis=-1)
