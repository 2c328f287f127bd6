#pragma once
#include <cuda_runtime.h>
#include "cuda.hpp"
namespace cv { namespace cuda { struct StreamAccessor { static cudaStream_t getStream(const Stream &s); }; } }
