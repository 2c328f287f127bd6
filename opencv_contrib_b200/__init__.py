"""opencv_contrib_b200 -- B200-native (sm_100a) re-implementation of the dense optical-flow path
of opencv_contrib's cudaoptflow module (cv::cuda::DenseOpticalFlow family).

Only what the hot path needs lives here:
  csrc/            hand-written CUDA kernels + C++ host engines + the C ABI (libb200flow.so)
  _lib.py          ctypes binding of include/b200flow.h
  cudaoptflow.py   Python mirror of the reference's generated bindings
                   (cv2.cuda.OpticalFlowDual_TVL1_create(...).calc(I0, I1, flow, stream))
  batch.py         batched frame-pair front end, one process per GPU (torch.distributed / NCCL)
  video.py         consecutive-frame front end (one upload per frame, warm start, 3-stream pipeline)
  frames.py        interpolateFrames, the consumer right after the solvers
  flowio.py        Middlebury .flo files + the reference's error statistics
"""
from .cudaoptflow import (  # noqa: F401
    DenseOpticalFlow,
    OpticalFlowDual_TVL1,
    FarnebackOpticalFlow,
    BroxOpticalFlow,
    DensePyrLKOpticalFlow,
    OpticalFlowDual_TVL1_create,
    FarnebackOpticalFlow_create,
    BroxOpticalFlow_create,
    DensePyrLKOpticalFlow_create,
    SparsePyrLKOpticalFlow,
    SparsePyrLKOpticalFlow_create,
    OPTFLOW_USE_INITIAL_FLOW,
    OPTFLOW_FARNEBACK_GAUSSIAN,
)
from ._lib import B2FError  # noqa: F401
from .frames import interpolateFrames  # noqa: F401
from .video import VideoFlow  # noqa: F401
from . import flowio  # noqa: F401
