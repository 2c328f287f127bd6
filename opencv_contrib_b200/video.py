"""Video front end: flow between consecutive frames of one stream.

    vf = VideoFlow(alg, rows, cols, dtype=np.uint8, depth=3, warm_start=False)
    for frame in frames:                 # host numpy arrays
        p = vf.push(frame)               # index of the pair this frame completed, or -1 for the first
        if p >= 1: flow = vf.fetch(p - 1)   # fetch lags the push, so copies and solves overlap

The work is done by ``b2f_video_*`` in libb200flow.so (csrc/video.cu): each frame is uploaded once, the
solve of pair k overlaps the upload of frame k+1 and the download of flow k-1, and with ``warm_start``
pair k starts from flow k-1 (tvl1flow.cpp:203-207, farneback.cpp:179-188; the chaining the reference
test does by hand, test_optflow.cpp:328-334).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import B2F_8UC1, B2F_32FC1, B2FError


class VideoFlow:
    def __init__(self, alg, rows: int, cols: int, dtype=np.uint8, depth: int = 3, warm_start: bool = False):
        self._lib = _lib.lib()
        self._alg = alg  # keeps the engine handle alive
        self.rows, self.cols, self.depth = rows, cols, depth
        self.dtype = np.dtype(dtype)
        if self.dtype == np.uint8:
            t = B2F_8UC1
        elif self.dtype == np.float32:
            t = B2F_32FC1
        else:
            raise B2FError(2)
        self._v = C.c_void_p()
        st = self._lib.b2f_video_create(alg._h, rows, cols, t, depth, int(warm_start), C.byref(self._v))
        if st != 0:
            self._v = None
            raise B2FError(st)

    def close(self):
        v, self._v = getattr(self, "_v", None), None
        if v:
            self._lib.b2f_video_destroy(v)

    __del__ = close

    def push(self, frame: np.ndarray) -> int:
        if frame.dtype != self.dtype or frame.shape != (self.rows, self.cols) or frame.strides[1] != frame.itemsize:
            raise B2FError(3)
        idx = C.c_int64(-1)
        st = self._lib.b2f_video_push(self._v, frame.ctypes.data, frame.strides[0], C.byref(idx))
        if st != 0:
            raise B2FError(st)
        return int(idx.value)

    def fetch(self, pair_index: int, out: np.ndarray | None = None, copy: bool = True) -> np.ndarray:
        """Wait for a pair's flow.  ``copy=False`` returns a read-only view of the front end's pinned result
        ring (no 16 MB host copy at 1080p); it stays valid until ``depth`` further frames have been pushed."""
        if not copy:
            ptr, step = C.POINTER(C.c_float)(), C.c_size_t()
            st = self._lib.b2f_video_fetch_view(self._v, pair_index, C.byref(ptr), C.byref(step))
            if st != 0:
                raise B2FError(st)
            a = np.ctypeslib.as_array(ptr, shape=(self.rows, self.cols, 2))
            a.flags.writeable = False
            return a
        if out is None:
            out = np.empty((self.rows, self.cols, 2), np.float32)
        st = self._lib.b2f_video_fetch(self._v, pair_index, out.ctypes.data, out.strides[0])
        if st != 0:
            raise B2FError(st)
        return out

    def run(self, frames, copy: bool = True):
        """Generator: yields (pair_index, flow) for every consecutive pair of ``frames``, keeping
        ``depth - 1`` pairs in flight.  With ``copy=False`` each yielded flow is a view that is only valid
        until the generator is advanced again."""
        pending = []
        for f in frames:
            p = self.push(f)
            if p >= 0:
                pending.append(p)
            while len(pending) >= self.depth:
                q = pending.pop(0)
                yield q, self.fetch(q, copy=copy)
        for q in pending:
            yield q, self.fetch(q, copy=copy)
