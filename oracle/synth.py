"""Seeded synthetic frame pairs with known ground-truth flow.

Test infrastructure (see oracle/__init__.py).  No opencv_extra images exist in
the build container or on the GPU box, so every parity test runs on these.

Convention (reference: modules/cudaoptflow/src/cuda/tvl1flow.cu:119-120,
farneback.cu:167-168):  I0(x, y) ~= I1(x + u(x, y), y + v(x, y)).
We therefore draw a texture T, set I1 = T and I0 = T sampled at (x+u, y+v).
"""
from __future__ import annotations

import numpy as np
import cv2


def texture(h: int, w: int, seed: int, cell: int = 8, sigma: float = 1.5) -> np.ndarray:
    """Low-pass random texture in [0, 255] float32: uniform noise at 1/cell
    resolution, bicubic upsample, Gaussian blur, contrast stretch."""
    rng = np.random.default_rng(seed)
    gh, gw = (h + cell - 1) // cell + 2, (w + cell - 1) // cell + 2
    coarse = rng.random((gh, gw), dtype=np.float32)
    # add a second octave so coarse pyramid levels still have structure
    gh2, gw2 = (gh + 3) // 4 + 2, (gw + 3) // 4 + 2
    coarse2 = rng.random((gh2, gw2), dtype=np.float32)
    up = cv2.resize(coarse, (gw * cell, gh * cell), interpolation=cv2.INTER_CUBIC)
    up2 = cv2.resize(coarse2, (gw * cell, gh * cell), interpolation=cv2.INTER_CUBIC)
    img = (up + 1.5 * up2)[cell:cell + h, cell:cell + w]
    img = cv2.GaussianBlur(img, (0, 0), sigma)
    lo, hi = float(img.min()), float(img.max())
    return ((img - lo) * (255.0 / max(hi - lo, 1e-6))).astype(np.float32)


def flow_field(h: int, w: int, kind: str, seed: int = 0, mag: float = 6.0) -> np.ndarray:
    """Ground-truth flow (h, w, 2) float32.  kinds: const | affine | smooth | zero."""
    ys, xs = np.mgrid[0:h, 0:w].astype(np.float32)
    if kind == "zero":
        u = np.zeros((h, w), np.float32)
        v = np.zeros((h, w), np.float32)
    elif kind == "const":
        u = np.full((h, w), 2.5, np.float32)
        v = np.full((h, w), -1.25, np.float32)
    elif kind == "affine":
        th = np.deg2rad(0.5)
        s = 1.01
        cx, cy = (w - 1) / 2.0, (h - 1) / 2.0
        xr = s * (np.cos(th) * (xs - cx) - np.sin(th) * (ys - cy)) + cx
        yr = s * (np.sin(th) * (xs - cx) + np.cos(th) * (ys - cy)) + cy
        u = (xr - xs).astype(np.float32)
        v = (yr - ys).astype(np.float32)
    elif kind == "smooth":
        rng = np.random.default_rng(1000 + seed)
        gh, gw = max(h // 96, 2) + 2, max(w // 96, 2) + 2
        cu = rng.uniform(-1, 1, (gh, gw)).astype(np.float32)
        cv_ = rng.uniform(-1, 1, (gh, gw)).astype(np.float32)
        u = cv2.resize(cu, (w, h), interpolation=cv2.INTER_CUBIC)
        v = cv2.resize(cv_, (w, h), interpolation=cv2.INTER_CUBIC)
        n = np.sqrt(u * u + v * v).max()
        u = (u * (mag / max(n, 1e-6))).astype(np.float32)
        v = (v * (mag / max(n, 1e-6))).astype(np.float32)
    else:
        raise ValueError(kind)
    return np.stack([u, v], axis=-1)


def make_pair(h: int, w: int, seed: int = 0, kind: str = "const", dtype: str = "u8",
              mag: float = 6.0):
    """Return (I0, I1, gt_flow).  dtype 'u8' -> uint8 0..255; 'f32' -> float32 in [0,1]
    (Brox / float TV-L1 convention, reference test_optflow.cpp:79-84)."""
    pad = 16
    T = texture(h + 2 * pad, w + 2 * pad, seed)
    gt = flow_field(h, w, kind, seed, mag)
    ys, xs = np.mgrid[0:h, 0:w].astype(np.float32)
    mapx = xs + gt[..., 0] + pad
    mapy = ys + gt[..., 1] + pad
    I1 = T[pad:pad + h, pad:pad + w].copy()
    I0 = cv2.remap(T, mapx, mapy, cv2.INTER_CUBIC, borderMode=cv2.BORDER_REFLECT101)
    if dtype == "u8":
        I0 = np.clip(np.rint(I0), 0, 255).astype(np.uint8)
        I1 = np.clip(np.rint(I1), 0, 255).astype(np.uint8)
    elif dtype == "f32":
        I0 = (np.clip(I0, 0, 255) / 255.0).astype(np.float32)
        I1 = (np.clip(I1, 0, 255) / 255.0).astype(np.float32)
    else:
        raise ValueError(dtype)
    return I0, I1, gt
