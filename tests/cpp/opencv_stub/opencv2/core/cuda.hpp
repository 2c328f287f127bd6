#pragma once
#include <cstddef>
#include "../core.hpp"
namespace cv { namespace cuda {
class Stream {
public:
    static Stream &Null();
};
class GpuMat {
public:
    int flags, rows, cols;
    size_t step;
    unsigned char *data;
    GpuMat();
    int type() const;
    Size size() const;
    void create(int rows, int cols, int type);
};
} }
