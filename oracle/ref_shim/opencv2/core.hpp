// Minimal stand-in for the parts of opencv2/core the reference's CPU Dual TV-L1
// (/root/reference/modules/optflow/src/tvl1flow.cpp) uses, so that the UNMODIFIED reference source
// compiles into oracle/_ref/libtvl1_ref.so without an OpenCV build.  TEST INFRASTRUCTURE.
//
// Only the semantics that file relies on are provided: reference-counted float/byte matrices with
// ROI headers, row pointers, Size/Rect/Range/Scalar, parallel_for_, Ptr/makePtr, the assertion
// macros.  The three imgproc primitives it calls (resize, remap, medianBlur) are declared in
// opencv2/imgproc.hpp and implemented in ../ref_glue.cpp on top of the cv2-pinned C restatements
// of oracle/tvl1_cpu.c.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#define CV_OVERRIDE override
#define CV_EXPORTS
#define CV_EXPORTS_W
#define CV_WRAP
#define CV_INSTRUMENT_REGION()
#define CV_OCL_RUN(condition, func)
#define CV_Assert(expr) do { if (!(expr)) throw std::runtime_error(std::string("CV_Assert failed: ") + #expr); } while (0)
#define CV_DbgAssert(expr) assert(expr)

#define CV_8U 0
#define CV_32F 5
#define CV_CN_SHIFT 3
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn) - 1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC2 CV_MAKETYPE(CV_32F, 2)

namespace cv {

typedef unsigned char uchar;
typedef std::string String;
template <typename T> using Ptr = std::shared_ptr<T>;
template <typename T, typename... A> Ptr<T> makePtr(A &&...a) { return std::make_shared<T>(std::forward<A>(a)...); }

struct Size {
    int width = 0, height = 0;
    Size() {}
    Size(int w, int h) : width(w), height(h) {}
    int area() const { return width * height; }
    bool operator==(const Size &o) const { return width == o.width && height == o.height; }
    bool operator!=(const Size &o) const { return !(*this == o); }
};
struct Rect {
    int x, y, width, height;
    Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {}
};
struct Range {
    int start, end;
    Range(int s, int e) : start(s), end(e) {}
};
struct Scalar {
    double val[4];
    static Scalar all(double v) { Scalar s; s.val[0] = s.val[1] = s.val[2] = s.val[3] = v; return s; }
    double operator[](int i) const { return val[i]; }
};

// Dense 2-D matrix, shared storage, byte step; type() uses OpenCV's numeric flags.
class Mat {
public:
    int rows = 0, cols = 0;
    uchar *data = nullptr;
    size_t step = 0;  // bytes per row

    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    void create(int r, int c, int type) {
        if (data && r == rows && c == cols && type == type_) return;
        rows = r; cols = c; type_ = type;
        step = (size_t)c * elemSize();
        store_.reset(new uchar[(size_t)r * step + 64], std::default_delete<uchar[]>());  // uninitialised, like cv::Mat
        data = store_.get();
        first_touch(data, r, step);
    }
    void create(Size s, int type) { create(s.height, s.width, type); }
    int type() const { return type_; }
    int depth() const { return type_ & 7; }
    int channels() const { return (type_ >> CV_CN_SHIFT) + 1; }
    size_t elemSize() const { return (size_t)channels() * (depth() == CV_8U ? 1 : 4); }
    Size size() const { return Size(cols, rows); }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    template <typename T> T *ptr(int y) { return reinterpret_cast<T *>(data + (ptrdiff_t)y * (ptrdiff_t)step); }
    template <typename T> const T *ptr(int y) const { return reinterpret_cast<const T *>(data + (ptrdiff_t)y * (ptrdiff_t)step); }
    Mat roi(const Rect &r) const {
        Mat m(*this);
        m.data = data + (size_t)r.y * step + (size_t)r.x * elemSize();
        m.rows = r.height; m.cols = r.width;
        return m;
    }
    void release() { *this = Mat(); }
    // dst = saturate_cast<float>(src * alpha), float arithmetic (cvtScale 8u->32f / 32f->32f)
    void convertTo(Mat &dst, int rtype, double alpha = 1.0) const;
    void setTo(const Scalar &s);

protected:
    int type_ = 0;
    std::shared_ptr<uchar> store_;
    // pages are faulted in by the OpenMP threads that will work on those rows (ref_glue.cpp: parallel_for_ uses the same
    // static row partition), so the timed CPU baseline is NUMA-local like an OpenCV build with its own thread pool
    static void first_touch(uchar *p, int rows, size_t step);
};

template <typename T> class Mat_;
template <> class Mat_<float> : public Mat {
public:
    Mat_() { type_ = CV_32FC1; }
    Mat_(const Mat &m) : Mat(m) { if (!m.empty()) CV_Assert(m.type() == CV_32FC1); type_ = CV_32FC1; }
    void create(Size s) { Mat::create(s.height, s.width, CV_32FC1); }
    void create(int r, int c) { Mat::create(r, c, CV_32FC1); }
    float *operator[](int y) { return ptr<float>(y); }
    const float *operator[](int y) const { return ptr<float>(y); }
    float &operator()(int y, int x) { return ptr<float>(y)[x]; }
    const float &operator()(int y, int x) const { return ptr<float>(y)[x]; }
    Mat_<float> operator()(const Rect &r) const { return Mat_<float>(roi(r)); }
};

inline void Mat::convertTo(Mat &dst, int rtype, double alpha) const {
    CV_Assert((rtype & 7) == CV_32F && channels() == 1);
    Mat out;
    out.create(rows, cols, CV_32FC1);
    const float a = (float)alpha;
    for (int y = 0; y < rows; ++y) {
        float *d = out.ptr<float>(y);
        if (depth() == CV_8U) { const uchar *s = ptr<uchar>(y); for (int x = 0; x < cols; ++x) d[x] = (float)s[x] * a; }
        else { const float *s = ptr<float>(y); for (int x = 0; x < cols; ++x) d[x] = s[x] * a; }
    }
    static_cast<Mat &>(dst) = out;
}
inline void Mat::setTo(const Scalar &s) {
    CV_Assert(depth() == CV_32F);
    const float v = (float)s[0];
    for (int y = 0; y < rows; ++y) { float *d = ptr<float>(y); for (int x = 0; x < cols * channels(); ++x) d[x] = v; }
}

// proxy argument types: the reference only asks them for getMat / size / type / isUMat
class _InputArray {
public:
    _InputArray(Mat &m) : m_(&m) {}
    Mat getMat() const { return *m_; }
    Size size() const { return m_->size(); }
    int type() const { return m_->type(); }
    bool isUMat() const { return false; }
    Mat &ref() const { return *m_; }
protected:
    Mat *m_;
};
class _InputOutputArray : public _InputArray {
public:
    _InputOutputArray(Mat &m) : _InputArray(m) {}
};
typedef const _InputArray &InputArray;
typedef const _InputOutputArray &InputOutputArray;
typedef const _InputOutputArray &OutputArray;

class Algorithm {
public:
    virtual ~Algorithm() {}
    virtual String getDefaultName() const { return "my_object"; }
};

struct ParallelLoopBody {
    virtual ~ParallelLoopBody() {}
    virtual void operator()(const Range &range) const = 0;
};
void parallel_for_(const Range &range, const ParallelLoopBody &body, double nstripes = -1.);

// cv::multiply(src, Scalar, dst): for CV_32F data the scalar is applied in float
void multiply(const Mat_<float> &src, const Scalar &s, Mat_<float> &dst);
void split(const Mat &src, Mat_<float> *mv);
void merge(const Mat *mv, size_t count, InputOutputArray dst);

}  // namespace cv
