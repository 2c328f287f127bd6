// flowio.cu -- host side of SURVEY.md 8f rank 3: Middlebury .flo files and the error measures the
// reference's tests and evaluation sample use.  No device code; lives in libb200flow.so so that the
// C++ adapter, the ctypes tests and tools/flow_eval.py share one implementation.
//
//  * format: optflow/test/test_tvl1optflow.cpp:49-108 (tag 202021.25f == "PIEH", int32 w, int32 h, then
//    row-major interleaved float32 (u, v));
//  * validity / endpoint / angular error, R and A statistics: optflow/samples/optical_flow_evaluation.cpp:23-163;
//  * regression criterion: optflow/test/test_tvl1optflow.cpp:114-142.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <vector>

#include "b200flow.h"

namespace {

const float FLO_TAG_FLOAT = 202021.25f;

struct File {
    FILE *f;
    explicit File(const char *path, const char *mode) : f(path ? std::fopen(path, mode) : nullptr) {}
    ~File() { if (f) std::fclose(f); }
};

inline bool flow_ok(float x, float y) {  // isFlowCorrect, optical_flow_evaluation.cpp:23-26
    return !std::isnan(x) && !std::isnan(y) && std::fabs(x) < 1e9 && std::fabs(y) < 1e9;
}

inline const float *rowc(const float *p, size_t step, int y) {
    return reinterpret_cast<const float *>(reinterpret_cast<const char *>(p) + step * (size_t)y);
}
inline float *row(float *p, size_t step, int y) {
    return reinterpret_cast<float *>(reinterpret_cast<char *>(p) + step * (size_t)y);
}

int read_header(FILE *f, int *rows, int *cols) {
    float tag = 0.f;
    int32_t w = 0, h = 0;
    if (std::fread(&tag, sizeof(float), 1, f) != 1) return B2F_BAD_ARG;
    if (tag != FLO_TAG_FLOAT) return B2F_UNSUPPORTED_TYPE;  // CV_Assert(tag == FLO_TAG_FLOAT), :88
    if (std::fread(&w, sizeof(int32_t), 1, f) != 1 || std::fread(&h, sizeof(int32_t), 1, f) != 1) return B2F_BAD_ARG;
    if (w <= 0 || h <= 0 || w > (1 << 20) || h > (1 << 20)) return B2F_BAD_ARG;
    *rows = h;
    *cols = w;
    return B2F_OK;
}

}  // namespace

extern "C" {

int b2f_flo_read_size(const char *path, int *rows, int *cols) {
    if (!path || !rows || !cols) return B2F_BAD_ARG;
    File file(path, "rb");
    if (!file.f) return B2F_BAD_ARG;
    return read_header(file.f, rows, cols);
}

int b2f_flo_read(const char *path, float *flow, size_t step, int rows, int cols) {
    if (!path || !flow) return B2F_BAD_ARG;
    File file(path, "rb");
    if (!file.f) return B2F_BAD_ARG;
    int r = 0, c = 0;
    const int st = read_header(file.f, &r, &c);
    if (st != B2F_OK) return st;
    if (r != rows || c != cols) return B2F_SIZE_MISMATCH;
    if (step < sizeof(float) * 2 * (size_t)cols) return B2F_BAD_ARG;
    for (int y = 0; y < rows; ++y)
        if (std::fread(row(flow, step, y), sizeof(float) * 2, (size_t)cols, file.f) != (size_t)cols) return B2F_BAD_ARG;
    return B2F_OK;
}

int b2f_flo_write(const char *path, const float *flow, size_t step, int rows, int cols) {
    if (!path || !flow || rows <= 0 || cols <= 0) return B2F_BAD_ARG;
    if (step < sizeof(float) * 2 * (size_t)cols) return B2F_BAD_ARG;
    File file(path, "wb");
    if (!file.f) return B2F_BAD_ARG;
    const int32_t w = cols, h = rows;
    if (std::fwrite("PIEH", 1, 4, file.f) != 4) return B2F_BAD_ARG;  // FLO_TAG_STRING, test_tvl1optflow.cpp:56
    if (std::fwrite(&w, sizeof(int32_t), 1, file.f) != 1 || std::fwrite(&h, sizeof(int32_t), 1, file.f) != 1)
        return B2F_BAD_ARG;
    for (int y = 0; y < rows; ++y)
        if (std::fwrite(rowc(flow, step, y), sizeof(float) * 2, (size_t)cols, file.f) != (size_t)cols) return B2F_BAD_ARG;
    return B2F_OK;
}

int b2f_flow_error_map(const float *flow1, size_t step1, const float *flow2, size_t step2, int rows, int cols,
                       int measure, float *err, size_t err_step) {
    if (!flow1 || !flow2 || !err || rows <= 0 || cols <= 0) return B2F_BAD_ARG;
    if (measure != B2F_ERR_ENDPOINT && measure != B2F_ERR_ANGULAR_REFERENCE && measure != B2F_ERR_ANGULAR)
        return B2F_BAD_ARG;
    const float nan = std::numeric_limits<float>::quiet_NaN();
    for (int y = 0; y < rows; ++y) {
        const float *a = rowc(flow1, step1, y), *b = rowc(flow2, step2, y);
        float *e = row(err, err_step, y);
        for (int x = 0; x < cols; ++x) {
            const float ax = a[2 * x], ay = a[2 * x + 1], bx = b[2 * x], by = b[2 * x + 1];
            if (!flow_ok(ax, ay) || !flow_ok(bx, by)) {
                e[x] = nan;
                continue;
            }
            if (measure == B2F_ERR_ENDPOINT) {
                // Point2f diff; diff.ddot(diff) in double; sqrt((float)...) (:41-42)
                const float dx = ax - bx, dy = ay - by;
                const double dd = (double)dx * dx + (double)dy * dy;
                e[x] = std::sqrt((float)dd);
            } else {
                // Point3f (u, v, 1); ddot and norm in double (:62-67)
                const double dot = (double)ax * bx + (double)ay * by + 1.0;
                const double n1 = std::sqrt((double)ax * ax + (double)ay * ay + 1.0);
                const double n2 = std::sqrt((double)bx * bx + (double)by * by + 1.0);
                const double arg = measure == B2F_ERR_ANGULAR_REFERENCE ? dot / n1 * n2 : dot / (n1 * n2);
                e[x] = std::acos((float)arg);
            }
        }
    }
    return B2F_OK;
}

int b2f_flow_error_stats(const float *err, size_t err_step, const unsigned char *mask, size_t mask_step, int rows,
                         int cols, b2f_error_stats *out) {
    if (!err || !out || rows <= 0 || cols <= 0) return B2F_BAD_ARG;
    std::memset(out, 0, sizeof(*out));
    static const float R_thresholds[5] = {0.5f, 1.f, 2.f, 5.f, 10.f};
    static const float A_thresholds[3] = {0.5f, 0.75f, 0.95f};
    auto masked = [&](int y, int x) { return !mask || mask[(size_t)y * mask_step + x] != 0; };

    // meanStdDev over the mask: sums in double; a NaN error propagates exactly as it does through cv::meanStdDev
    double sum = 0.0, sqsum = 0.0, maxv = -std::numeric_limits<double>::infinity();
    int64_t n = 0, rc[5] = {0, 0, 0, 0, 0};
    for (int y = 0; y < rows; ++y) {
        const float *e = rowc(err, err_step, y);
        for (int x = 0; x < cols; ++x) {
            if (!masked(y, x)) continue;
            ++n;
            const float v = e[x];
            sum += v;
            sqsum += (double)v * v;
            if (v > maxv) maxv = v;  // minMaxLoc ignores NaN (comparison false)
            for (int k = 0; k < 5; ++k) rc[k] += v > R_thresholds[k];
        }
    }
    out->count = n;
    if (n == 0) return B2F_OK;
    const double mean = sum / n;
    out->mean = mean;
    out->stddev = std::sqrt(std::max(sqsum / n - mean * mean, 0.0));
    out->max = maxv;
    for (int k = 0; k < 5; ++k) out->r[k] = (float)rc[k] / n;  // stat_RX returns (float)count / all

    // calcHist: 1024 uniform bins over [0, max); values outside (incl. == max and NaN) fall in no bin
    const int bins = 1024;
    std::vector<int64_t> hist(bins, 0);
    const float hi = (float)maxv;
    if (hi > 0.f) {
        const double a = bins / (double)hi;
        for (int y = 0; y < rows; ++y) {
            const float *e = rowc(err, err_step, y);
            for (int x = 0; x < cols; ++x) {
                if (!masked(y, x)) continue;
                const float v = e[x];
                if (!(v >= 0.f) || !(v < hi)) continue;
                int idx = (int)std::floor(v * a);
                if (idx >= bins) idx = bins - 1;
                hist[idx]++;
            }
        }
    }
    for (int k = 0; k < 3; ++k) {
        const int cutoff = (int)std::floor(A_thresholds[k] * n + 0.5f);
        int64_t counter = 0;
        int bin = 0;
        while (bin < bins && counter < cutoff) {  // stat_AX, :94-105
            counter += hist[bin];
            ++bin;
        }
        out->a[k] = (float)bin / bins * hi;
    }
    return B2F_OK;
}

int b2f_flow_accuracy(const float *gold, size_t gold_step, const float *flow, size_t flow_step, int rows, int cols,
                      double threshold, double *fraction) {
    if (!gold || !flow || !fraction || rows <= 0 || cols <= 0) return B2F_BAD_ARG;
    const double thr2 = threshold * threshold;
    size_t gold_counter = 0, valid_counter = 0;
    for (int y = 0; y < rows; ++y) {
        const float *g = rowc(gold, gold_step, y), *f = rowc(flow, flow_step, y);
        for (int x = 0; x < cols; ++x) {
            if (!flow_ok(g[2 * x], g[2 * x + 1])) continue;
            gold_counter++;
            if (!flow_ok(f[2 * x], f[2 * x + 1])) continue;
            const float dx = g[2 * x] - f[2 * x], dy = g[2 * x + 1] - f[2 * x + 1];
            const double e = (double)dx * dx + (double)dy * dy;
            if (e <= thr2) valid_counter++;
        }
    }
    *fraction = gold_counter ? (double)valid_counter / (double)gold_counter : 1.0;
    return B2F_OK;
}

}  // extern "C"
