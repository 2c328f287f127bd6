"""ctypes loader for oracle/tvl1_cpu.c (the C/OpenMP port of the reference's CPU TV-L1).
TEST INFRASTRUCTURE (see oracle/__init__.py): used by tests and by bench.py's cpu_baseline /
--impl reference legs only."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_build", "libtvl1_cpu.so")
_lib = None


class _Params(C.Structure):
    _fields_ = [("tau", C.c_double), ("lambda_", C.c_double), ("theta", C.c_double), ("nscales", C.c_int),
                ("warps", C.c_int), ("epsilon", C.c_double), ("innerIterations", C.c_int),
                ("outerIterations", C.c_int), ("scaleStep", C.c_double), ("gamma", C.c_double),
                ("medianFiltering", C.c_int), ("useInitialFlow", C.c_int)]


def available() -> bool:
    return os.path.exists(_PATH)


def usable_cpus() -> int:
    """CPUs this process may really use: affinity mask capped by the cgroup CPU quota (a container on a
    128-thread host is often limited to a few cores; 128 spinning OpenMP threads on 16 cores run ~35x slower)."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    n = min(n, max(1, int(int(txt[0]) / int(txt[1]) + 0.5)))
            else:
                q = int(txt[0])
                if q > 0:
                    per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
                    n = min(n, max(1, int(q / per + 0.5)))
            break
        except Exception:
            continue
    return max(1, n)


def set_threads(n: int) -> int:
    l = lib()
    return int(l.tvl1_cpu_set_threads(int(n)))


def lib():
    global _lib
    if _lib is None:
        os.environ.setdefault("OMP_WAIT_POLICY", "passive")
        _lib = C.CDLL(_PATH)
        _lib.tvl1_cpu_set_threads.restype = C.c_int
        _lib.tvl1_cpu_set_threads.argtypes = [C.c_int]
        _lib.tvl1_cpu_set_threads(usable_cpus())
        fp = C.POINTER(C.c_float)
        _lib.tvl1_cpu_calc.restype = C.c_int
        _lib.tvl1_cpu_calc.argtypes = [C.POINTER(_Params), fp, fp, C.c_int, C.c_int, fp]
        _lib.tvl1_cpu_resize_linear.argtypes = [fp, C.c_int, C.c_int, fp, C.c_int, C.c_int]
        _lib.tvl1_cpu_remap_cubic.argtypes = [fp, C.c_int, C.c_int, fp, fp, fp]
    return _lib


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def calc(I0: np.ndarray, I1: np.ndarray, P) -> np.ndarray:
    """P: oracle.tvl1_cpu.TVL1Params (gamma == 0, medianFiltering == 1, no initial flow)."""
    k = np.float32(1.0) if I0.dtype == np.uint8 else np.float32(255.0)
    a = np.ascontiguousarray(I0.astype(np.float32) * k)
    b = np.ascontiguousarray(I1.astype(np.float32) * k)
    h, w = a.shape
    flow = np.empty((h, w, 2), np.float32)
    p = _Params(P.tau, P.lambda_, P.theta, P.nscales, P.warps, P.epsilon, P.innerIterations, P.outerIterations,
                P.scaleStep, P.gamma, P.medianFiltering, int(P.useInitialFlow))
    rc = lib().tvl1_cpu_calc(C.byref(p), _fp(a), _fp(b), h, w, _fp(flow))
    if rc != 0:
        raise ValueError("tvl1_cpu_calc: unsupported parameters (gamma, medianFiltering>1 or initial flow)")
    return flow


def resize_linear(src: np.ndarray, dh: int, dw: int) -> np.ndarray:
    src = np.ascontiguousarray(src, np.float32)
    dst = np.empty((dh, dw), np.float32)
    lib().tvl1_cpu_resize_linear(_fp(src), src.shape[0], src.shape[1], _fp(dst), dh, dw)
    return dst


def remap_cubic(src: np.ndarray, mapx: np.ndarray, mapy: np.ndarray) -> np.ndarray:
    src = np.ascontiguousarray(src, np.float32)
    mapx = np.ascontiguousarray(mapx, np.float32)
    mapy = np.ascontiguousarray(mapy, np.float32)
    dst = np.empty(src.shape, np.float32)
    lib().tvl1_cpu_remap_cubic(_fp(src), src.shape[0], src.shape[1], _fp(mapx), _fp(mapy), _fp(dst))
    return dst
