// Declaration-only stand-in for the parts of OpenCV's core API that include/b200flow/cudaoptflow_compat.hpp touches in its
// REAL-OpenCV branch (-DB200FLOW_WITH_OPENCV).  Signatures follow opencv core 4.x (core/mat.hpp, core/cuda.hpp,
// core/cvstd_wrapper.hpp): in particular _InputArray::getGpuMat() returns BY VALUE and getGpuMatRef() by reference --
// the distinction that broke the branch in round 1.  Used only by the compile-only target `opencv_branch_syntax`.
#pragma once
#include <memory>
#include <string>
#define CV_32FC2 13
namespace cv {
typedef std::string String;
template <class T> using Ptr = std::shared_ptr<T>;
template <class T, class... A> Ptr<T> makePtr(A &&...a) { return std::make_shared<T>(std::forward<A>(a)...); }
struct Size {
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
};
namespace cuda { class GpuMat; class Stream; }
class _InputArray {
public:
    _InputArray(const cuda::GpuMat &m);
    cuda::GpuMat getGpuMat() const;
};
class _OutputArray : public _InputArray {
public:
    _OutputArray(cuda::GpuMat &m);
    void create(Size sz, int type) const;
    void create(int rows, int cols, int type) const;
    cuda::GpuMat &getGpuMatRef() const;
    bool needed() const;
};
class _InputOutputArray : public _OutputArray {
public:
    _InputOutputArray(cuda::GpuMat &m);
};
typedef const _InputArray &InputArray;
typedef const _OutputArray &OutputArray;
typedef const _InputOutputArray &InputOutputArray;
InputOutputArray noArray();
class Algorithm {
public:
    virtual ~Algorithm();
    virtual String getDefaultName() const;
};
namespace Error { enum Code { StsAssert = -215, GpuNotSupported = -216, GpuApiCallError = -217 }; }
[[noreturn]] void error(int code, const String &err, const char *func, const char *file, int line);
}  // namespace cv
