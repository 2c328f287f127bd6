// ref_glue.cpp -- what the UNMODIFIED reference source modules/optflow/src/tvl1flow.cpp needs around it to
// become oracle/_ref/libtvl1_ref.so (recipe: oracle/Makefile, target _ref/libtvl1_ref.so):
//   * the non-template parts of the opencv2 stand-in headers (ref_shim/opencv2/*.hpp);
//   * cv::resize / cv::remap / cv::medianBlur delegating to the C restatements of oracle/tvl1_cpu.c
//     (pinned against cv2 4.13 in tests/test_oracle_cpu.py) -- these three are external to
//     /root/reference (opencv/opencv imgproc), so a restatement is the best available here;
//   * a C entry point for ctypes.
// TEST INFRASTRUCTURE: never linked into libb200flow.so.
#include <opencv2/optflow.hpp>
#include <opencv2/imgproc.hpp>
#ifdef _OPENMP
#include <omp.h>
#endif

extern "C" {
void tvl1_cpu_resize_linear_f(const float *src, int sh, int sw, float *dst, int dh, int dw, double f);
void tvl1_cpu_remap_cubic(const float *src, int h, int w, const float *mapx, const float *mapy, float *dst);
void tvl1_cpu_median_blur(const float *src, int h, int w, float *dst, int ksize);
int tvl1_cpu_set_threads(int n);
}

namespace cv {

void parallel_for_(const Range &range, const ParallelLoopBody &body, double) {
    const int n = range.end - range.start;
    if (n <= 0) return;
#ifdef _OPENMP
    const int chunks = std::min(n, omp_get_max_threads() * 4);
#pragma omp parallel for schedule(static)
    for (int c = 0; c < chunks; ++c) {
        const int a = range.start + (int)((long long)n * c / chunks), b = range.start + (int)((long long)n * (c + 1) / chunks);
        if (b > a) body(Range(a, b));
    }
#else
    body(range);
#endif
}

void Mat::first_touch(uchar *p, int rows, size_t step) {
#pragma omp parallel for schedule(static)
    for (int y = 0; y < rows; ++y)
        for (size_t x = 0; x < step; x += 4096) p[(size_t)y * step + x] = 0;
}

// a contiguous copy of a (possibly ROI) float matrix
static std::vector<float> dense(const Mat_<float> &m) {
    std::vector<float> v((size_t)m.rows * m.cols);
    for (int y = 0; y < m.rows; ++y) std::memcpy(&v[(size_t)y * m.cols], m[y], sizeof(float) * m.cols);
    return v;
}
static void scatter(const std::vector<float> &v, Mat_<float> &m) {
    for (int y = 0; y < m.rows; ++y) std::memcpy(m[y], &v[(size_t)y * m.cols], sizeof(float) * m.cols);
}

void multiply(const Mat_<float> &src, const Scalar &s, Mat_<float> &dst) {
    const float k = (float)s[0];
    if (dst.size() != src.size()) dst.create(src.size());
    for (int y = 0; y < src.rows; ++y) {
        const float *a = src[y];
        float *d = dst[y];
        for (int x = 0; x < src.cols; ++x) d[x] = a[x] * k;
    }
}

void split(const Mat &src, Mat_<float> *mv) {
    CV_Assert(src.type() == CV_32FC2);
    for (int c = 0; c < 2; ++c)
        if (mv[c].size() != src.size()) mv[c].create(src.size());
    for (int y = 0; y < src.rows; ++y) {
        const float *s = src.ptr<float>(y);
        float *a = mv[0][y], *b = mv[1][y];
        for (int x = 0; x < src.cols; ++x) { a[x] = s[2 * x]; b[x] = s[2 * x + 1]; }
    }
}

void merge(const Mat *mv, size_t count, InputOutputArray dst) {
    CV_Assert(count == 2 && mv[0].type() == CV_32FC1 && mv[1].size() == mv[0].size());
    Mat &d = dst.ref();
    d.create(mv[0].rows, mv[0].cols, CV_32FC2);
    for (int y = 0; y < d.rows; ++y) {
        const float *a = mv[0].ptr<float>(y), *b = mv[1].ptr<float>(y);
        float *o = d.ptr<float>(y);
        for (int x = 0; x < d.cols; ++x) { o[2 * x] = a[x]; o[2 * x + 1] = b[x]; }
    }
}

static int cv_round_half_even(double v) { return (int)std::nearbyint(v); }

void resize(const Mat_<float> &src, Mat_<float> &dst, Size dsize, double fx, double fy, int interpolation) {
    CV_Assert(interpolation == INTER_LINEAR && !src.empty());
    double f = 0.0;
    if (dsize.area() == 0) {  // imgproc resize.cpp: dsize = Size(saturate_cast<int>(cols*fx), saturate_cast<int>(rows*fy))
        CV_Assert(fx > 0 && fx == fy);
        dsize = Size(cv_round_half_even(src.cols * fx), cv_round_half_even(src.rows * fy));
        f = fx;
    }
    const std::vector<float> s = dense(src);
    std::vector<float> d((size_t)dsize.area());
    tvl1_cpu_resize_linear_f(s.data(), src.rows, src.cols, d.data(), dsize.height, dsize.width, f);
    dst.create(dsize);
    scatter(d, dst);
}

void remap(const Mat_<float> &src, Mat_<float> &dst, const Mat_<float> &map1, const Mat_<float> &map2, int interpolation) {
    CV_Assert(interpolation == INTER_CUBIC && map1.size() == src.size() && map2.size() == src.size());
    const std::vector<float> s = dense(src), m1 = dense(map1), m2 = dense(map2);
    std::vector<float> d(s.size());
    tvl1_cpu_remap_cubic(s.data(), src.rows, src.cols, m1.data(), m2.data(), d.data());
    if (dst.size() != src.size()) dst.create(src.size());
    scatter(d, dst);
}

void medianBlur(const Mat_<float> &src, Mat_<float> &dst, int ksize) {
    CV_Assert(ksize == 3 || ksize == 5);  // the float path of cv::medianBlur accepts 3 and 5 only
    const std::vector<float> s = dense(src);
    std::vector<float> d(s.size());
    tvl1_cpu_median_blur(s.data(), src.rows, src.cols, d.data(), ksize);
    if (dst.size() != src.size()) dst.create(src.size());
    scatter(d, dst);
}

}  // namespace cv

struct tvl1_ref_params {
    double tau, lambda, theta;
    int nscales, warps;
    double epsilon;
    int innerIterations, outerIterations;
    double scaleStep, gamma;
    int medianFiltering, useInitialFlow;
};

#define REF_API extern "C" __attribute__((visibility("default")))

// I0/I1: rows x cols, CV_8UC1 (is_u8) or CV_32FC1; flow: rows x cols x 2 float32, read first when
// useInitialFlow is set.  Returns 0, or -1 with the CV_Assert text in err (when err != NULL).
REF_API int tvl1_ref_calc(const tvl1_ref_params *P, const void *I0, const void *I1, int is_u8, int rows, int cols,
                          float *flow, char *err, int errlen) {
    try {
        const int type = is_u8 ? CV_8UC1 : CV_32FC1;
        cv::Mat a(rows, cols, type), b(rows, cols, type), f;
        std::memcpy(a.data, I0, (size_t)rows * a.step);
        std::memcpy(b.data, I1, (size_t)rows * b.step);
        if (P->useInitialFlow) {
            f.create(rows, cols, CV_32FC2);
            std::memcpy(f.data, flow, (size_t)rows * f.step);
        }
        cv::Ptr<cv::optflow::DualTVL1OpticalFlow> alg = cv::optflow::DualTVL1OpticalFlow::create(
            P->tau, P->lambda, P->theta, P->nscales, P->warps, P->epsilon, P->innerIterations, P->outerIterations,
            P->scaleStep, P->gamma, P->medianFiltering, P->useInitialFlow != 0);
        alg->calc(cv::_InputArray(a), cv::_InputArray(b), cv::_InputOutputArray(f));
        CV_Assert(f.rows == rows && f.cols == cols && f.type() == CV_32FC2);
        for (int y = 0; y < rows; ++y) std::memcpy(flow + (size_t)y * cols * 2, f.ptr<float>(y), sizeof(float) * 2 * cols);
        return 0;
    } catch (const std::exception &e) {
        if (err && errlen > 0) { std::strncpy(err, e.what(), errlen - 1); err[errlen - 1] = 0; }
        return -1;
    }
}
REF_API int tvl1_ref_set_threads(int n) { return tvl1_cpu_set_threads(n); }
REF_API const char *tvl1_ref_source(void) { return TVL1_REF_SOURCE; }
