"""cv::cuda::interpolateFrames (cudalegacy.hpp:229; src/interpolate_frames.cpp:54-111).

    newFrame = interpolateFrames(frame0, frame1, fu, fv, bu, bv, pos[, newFrame[, buf[, stream]]])

All images are CUDA float32 H x W tensors sharing one row pitch (the reference asserts equal steps);
``buf`` is the 6H x W scratch the reference exposes.  ``corrected=False`` reproduces the reference
including its defects (see include/b200flow.h), ``corrected=True`` fixes them.
"""
from __future__ import annotations

import ctypes as C

from . import _lib
from ._lib import B2FError
from .cudaoptflow import _image_from_tensor, _torch

INTERP_REFERENCE, INTERP_CORRECTED = 0, 1


def interpolateFrames(frame0, frame1, fu, fv, bu, bv, pos: float, newFrame=None, buf=None, stream=None,
                      corrected: bool = False):
    torch = _torch()
    l = _lib.lib()
    if newFrame is None:
        newFrame = torch.empty_like(frame0)
    if buf is None:
        buf = torch.empty((6 * frame0.shape[0], frame0.shape[1]), dtype=torch.float32, device=frame0.device)
    imgs = [_image_from_tensor(t) for t in (frame0, frame1, fu, fv, bu, bv, newFrame, buf)]
    if stream is None:
        stream = torch.cuda.current_stream(frame0.device)
    sptr = stream.cuda_stream if hasattr(stream, "cuda_stream") else int(stream)
    st = l.b2f_interpolate_frames(*[C.byref(i) for i in imgs[:6]], C.c_float(pos), C.byref(imgs[6]),
                                  C.byref(imgs[7]), INTERP_CORRECTED if corrected else INTERP_REFERENCE,
                                  C.c_void_p(sptr))
    if st != 0:
        raise B2FError(st)
    return newFrame
