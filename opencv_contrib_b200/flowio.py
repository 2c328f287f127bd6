"""Middlebury .flo files and the reference's flow-error measures.

Thin ctypes wrappers over the host functions of libb200flow.so (csrc/flowio.cu); formats and
formulas follow optflow/test/test_tvl1optflow.cpp:49-142 and
optflow/samples/optical_flow_evaluation.cpp:23-163.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import B2FError, b2f_error_stats

ERR_ENDPOINT, ERR_ANGULAR_REFERENCE, ERR_ANGULAR = 0, 1, 2


def _flow(a: np.ndarray) -> np.ndarray:
    a = np.asarray(a)
    if a.ndim != 3 or a.shape[2] != 2 or a.dtype != np.float32 or a.strides[2] != 4 or a.strides[1] != 8:
        a = np.ascontiguousarray(a, np.float32)
        if a.ndim != 3 or a.shape[2] != 2:
            raise B2FError(2)
    return a


def readOpticalFlow(path: str) -> np.ndarray:
    """cv::readOpticalFlow / readOpticalFlowFromFile (test_tvl1optflow.cpp:80-108)."""
    l = _lib.lib()
    r, c = C.c_int(), C.c_int()
    st = l.b2f_flo_read_size(os.fsencode(path), C.byref(r), C.byref(c))
    if st != 0:
        raise B2FError(st)
    out = np.empty((r.value, c.value, 2), np.float32)
    st = l.b2f_flo_read(os.fsencode(path), out.ctypes.data, out.strides[0], r.value, c.value)
    if st != 0:
        raise B2FError(st)
    return out


def writeOpticalFlow(path: str, flow: np.ndarray) -> None:
    """cv::writeOpticalFlow / writeOpticalFlowToFile (test_tvl1optflow.cpp:52-76)."""
    f = _flow(flow)
    st = _lib.lib().b2f_flo_write(os.fsencode(path), f.ctypes.data, f.strides[0], f.shape[0], f.shape[1])
    if st != 0:
        raise B2FError(st)


def errorMap(flow1: np.ndarray, flow2: np.ndarray, measure: int = ERR_ENDPOINT) -> np.ndarray:
    a, b = _flow(flow1), _flow(flow2)
    if a.shape != b.shape:
        raise B2FError(3)
    err = np.empty(a.shape[:2], np.float32)
    st = _lib.lib().b2f_flow_error_map(a.ctypes.data, a.strides[0], b.ctypes.data, b.strides[0], a.shape[0],
                                       a.shape[1], measure, err.ctypes.data, err.strides[0])
    if st != 0:
        raise B2FError(st)
    return err


def errorStats(err: np.ndarray, mask: np.ndarray | None = None) -> dict:
    """calculateStats (optical_flow_evaluation.cpp:106-163): mean, std, R0.5..R10, A0.50..A0.95."""
    err = np.ascontiguousarray(err, np.float32)
    mptr, mstep = None, 0
    if mask is not None:
        mask = np.ascontiguousarray(mask, np.uint8)
        if mask.shape != err.shape:
            raise B2FError(3)
        mptr, mstep = mask.ctypes.data, mask.strides[0]
    s = b2f_error_stats()
    st = _lib.lib().b2f_flow_error_stats(err.ctypes.data, err.strides[0], mptr, mstep, err.shape[0], err.shape[1],
                                         C.byref(s))
    if st != 0:
        raise B2FError(st)
    return {"mean": s.mean, "std": s.stddev, "max": s.max, "count": int(s.count),
            "R": dict(zip((0.5, 1.0, 2.0, 5.0, 10.0), list(s.r))),
            "A": dict(zip((0.5, 0.75, 0.95), list(s.a)))}


def accuracy(gold: np.ndarray, flow: np.ndarray, threshold: float = 0.1) -> float:
    """Fraction of valid gold pixels with endpoint error <= threshold (test_tvl1optflow.cpp:114-142;
    the reference's regression test requires >= 0.95 at 0.1)."""
    g, f = _flow(gold), _flow(flow)
    if g.shape != f.shape:
        raise B2FError(3)
    out = C.c_double()
    st = _lib.lib().b2f_flow_accuracy(g.ctypes.data, g.strides[0], f.ctypes.data, f.strides[0], g.shape[0],
                                      g.shape[1], threshold, C.byref(out))
    if st != 0:
        raise B2FError(st)
    return out.value
