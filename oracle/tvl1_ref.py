"""ctypes loader for oracle/_ref/libtvl1_ref.so: the reference's OWN CPU Dual TV-L1
(/root/reference/modules/optflow/src/tvl1flow.cpp, compiled unmodified by oracle/Makefile against the
stand-in headers in oracle/ref_shim/; cv::resize / cv::remap / cv::medianBlur are the cv2-pinned C
restatements of oracle/tvl1_cpu.c).  This is what pins oracle/tvl1_cpu.py and oracle/tvl1_cpu.c
end to end: same inputs, the reference's own arithmetic.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The .so is built in the container that has
/root/reference and travels to the GPU box with the snapshot (git-ignored, not gpurun-ignored)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .tvl1_cpu_native import _Params, usable_cpus

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libtvl1_ref.so")
_lib = None


def available() -> bool:
    return os.path.exists(_PATH)


def lib():
    global _lib
    if _lib is None:
        os.environ.setdefault("OMP_WAIT_POLICY", "passive")
        _lib = C.CDLL(_PATH)
        _lib.tvl1_ref_calc.restype = C.c_int
        _lib.tvl1_ref_calc.argtypes = [C.POINTER(_Params), C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                       C.POINTER(C.c_float), C.c_char_p, C.c_int]
        _lib.tvl1_ref_set_threads.restype = C.c_int
        _lib.tvl1_ref_set_threads.argtypes = [C.c_int]
        _lib.tvl1_ref_source.restype = C.c_char_p
        _lib.tvl1_ref_set_threads(usable_cpus())
    return _lib


def set_threads(n: int) -> int:
    return int(lib().tvl1_ref_set_threads(int(n)))


def source() -> str:
    return lib().tvl1_ref_source().decode()


def calc(I0: np.ndarray, I1: np.ndarray, P, flow: np.ndarray | None = None) -> np.ndarray:
    """P: oracle.tvl1_cpu.TVL1Params (every field is honoured: gamma, medianFiltering, useInitialFlow)."""
    assert I0.dtype == I1.dtype and I0.dtype in (np.uint8, np.float32) and I0.shape == I1.shape
    a, b = np.ascontiguousarray(I0), np.ascontiguousarray(I1)
    h, w = a.shape
    out = np.zeros((h, w, 2), np.float32) if flow is None else np.ascontiguousarray(flow, np.float32).copy()
    p = _Params(P.tau, P.lambda_, P.theta, P.nscales, P.warps, P.epsilon, P.innerIterations, P.outerIterations,
                P.scaleStep, P.gamma, P.medianFiltering, int(P.useInitialFlow))
    err = C.create_string_buffer(512)
    rc = lib().tvl1_ref_calc(C.byref(p), a.ctypes.data, b.ctypes.data, int(a.dtype == np.uint8), h, w,
                             out.ctypes.data_as(C.POINTER(C.c_float)), err, 512)
    if rc != 0:
        raise ValueError("reference DualTVL1OpticalFlow::calc failed: " + err.value.decode())
    return out
