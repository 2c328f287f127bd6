"""Flow comparison metrics used by the parity tests (test infrastructure).

* endpoint error statistics — the criterion of the reference's CPU TV-L1
  regression test (modules/optflow/test/test_tvl1optflow.cpp:114-142:
  ">= 95 % of valid pixels with EPE <= 0.1").
* NCC dissimilarity — what the reference's EXPECT_MAT_SIMILAR computes
  (modules/cudaoptflow/test/test_optflow.cpp:341-348,462-465; implemented in
  opencv/opencv modules/ts as |1 - matchTemplate(TM_CCORR_NORMED)|).
"""
from __future__ import annotations

import numpy as np


def epe(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    d = a.astype(np.float64) - b.astype(np.float64)
    return np.sqrt((d * d).sum(axis=-1))


def epe_stats(a: np.ndarray, b: np.ndarray, border: int = 0) -> dict:
    e = epe(a, b)
    if border:
        e = e[border:-border, border:-border]
    return {
        "mean": float(e.mean()),
        "max": float(e.max()),
        "p95": float(np.percentile(e, 95)),
        "frac_le_0.1": float((e <= 0.1).mean()),
    }


def ncc_dissimilarity(a: np.ndarray, b: np.ndarray) -> float:
    """|1 - sum(a*b)/sqrt(sum(a^2) sum(b^2))| over all channels (TM_CCORR_NORMED)."""
    a = a.astype(np.float64).ravel()
    b = b.astype(np.float64).ravel()
    den = np.sqrt((a * a).sum() * (b * b).sum())
    if den == 0:
        return 0.0 if (a == b).all() else 1.0
    return float(abs(1.0 - (a * b).sum() / den))
