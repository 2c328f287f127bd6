/*
 * tvl1_cpu.c -- plain-C (OpenMP) restatement of the reference's CPU Dual TV-L1,
 * cv::optflow::DualTVL1OpticalFlow (modules/optflow/src/tvl1flow.cpp), for use as the timed CPU
 * baseline (bench.py cpu_baseline / --impl reference) and as a second oracle.
 *
 * TEST INFRASTRUCTURE -- never linked into libb200flow.so.
 *
 * The in-tree arithmetic follows tvl1flow.cpp line by line (citations below).  The three external
 * OpenCV primitives the reference calls are restated from their published algorithms and pinned
 * against the live cv2 4.13 functions in tests/test_oracle_cpu.py:
 *   cv::resize(INTER_LINEAR, float)  centre-aligned bilinear, float weights (imgproc resize.cpp)
 *   cv::remap(INTER_CUBIC, float)    a = -0.75 bicubic, coordinates quantised to 1/32 px,
 *                                    BORDER_CONSTANT 0 (imgproc imgwarp.cpp, INTER_TAB_SIZE = 32)
 *   cv::medianBlur(float, 3|5)       exact median of the k x k window, BORDER_REPLICATE (imgproc
 *                                    median_blur.simd.hpp, medianBlur_SortNet)
 * tvl1_cpu_calc itself implements medianFiltering == 1 (off), the setting the reference's own
 * GPU-vs-CPU test uses (modules/cudaoptflow/test/test_optflow.cpp:456-460); tvl1_cpu_median_blur
 * serves oracle/_ref (the reference's own source compiled against ref_shim/), which calls it.
 * Parallelism mirrors the reference: cv::parallel_for_ over rows in every stage
 * (tvl1flow.cpp:682,735,821,887,968,1068,1220) -> `omp parallel for` over rows; estimateU's error
 * accumulation is serial in the reference (:1087-1113) and is a per-row reduction here.
 */
#include <float.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include <stdlib.h>
#include <string.h>

typedef struct {
    double tau, lambda, theta;
    int nscales, warps;
    double epsilon;
    int innerIterations, outerIterations;
    double scaleStep, gamma;
    int medianFiltering, useInitialFlow;
} tvl1_cpu_params;

static int cv_round(double v) { return (int)nearbyint(v); }

/* malloc + parallel first touch: every stage below runs `omp parallel for schedule(static)` over rows, so the
 * thread that will work on a row is the one that faults its pages in (NUMA-local on a multi-socket host; without
 * this the main thread's memset/memcpy puts every page on one node and the timed baseline swings by 2x). */
static float *alloc_rows(int h, int w, int zero) {
    float *p = (float *)malloc(sizeof(float) * (size_t)h * w);
    if (!p) return NULL;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++) {
        float *r = p + (size_t)y * w;
        if (zero) memset(r, 0, sizeof(float) * (size_t)w);
        else for (int x = 0; x < w; x += 1024) r[x] = 0.f; /* touch each page */
    }
    return p;
}
static void copy_rows(float *dst, const float *src, int h, int w) {
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++) memcpy(dst + (size_t)y * w, src + (size_t)y * w, sizeof(float) * (size_t)w);
}

/* ---- cv::resize INTER_LINEAR, CV_32FC1 ------------------------------------------------------ */
/* f > 0: resize(src, dst, Size(), f, f) -- the scale is the given factor (pyramid, tvl1flow.cpp:479-480);
 * f <= 0: resize(src, dst, dsize)      -- the scale is dsize / ssize (flow prolongation, :520-522). */
static void resize_linear(const float *src, int sh, int sw, float *dst, int dh, int dw, double f) {
    if (dh == sh && dw == sw) {
        memcpy(dst, src, sizeof(float) * (size_t)sh * sw);
        return;
    }
    const double inv_x = f > 0 ? f : (double)dw / sw, inv_y = f > 0 ? f : (double)dh / sh;
    const double scale_x = 1. / inv_x, scale_y = 1. / inv_y;
    int *xofs = (int *)malloc(sizeof(int) * dw);
    float *alpha = (float *)malloc(sizeof(float) * 2 * dw);
    for (int dx = 0; dx < dw; dx++) {
        const double fxd = (dx + 0.5) * scale_x - 0.5; /* fraction formed in double (cv2 4.13 behaviour) */
        int sx = (int)floor(fxd);
        float fx = (float)(fxd - sx);
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        xofs[dx] = sx;
        alpha[2 * dx] = 1.f - fx;
        alpha[2 * dx + 1] = fx;
    }
#pragma omp parallel for schedule(static)
    for (int dy = 0; dy < dh; dy++) {
        const double fyd = (dy + 0.5) * scale_y - 0.5;
        int sy = (int)floor(fyd);
        float fy = (float)(fyd - sy);
        int sy0 = sy, sy1 = sy + 1;
        if (sy0 < 0) sy0 = 0;
        if (sy0 > sh - 1) sy0 = sh - 1;
        if (sy1 < 0) sy1 = 0;
        if (sy1 > sh - 1) sy1 = sh - 1;
        const float b0 = 1.f - fy, b1 = fy;
        const float *S0 = src + (size_t)sy0 * sw, *S1 = src + (size_t)sy1 * sw;
        float *D = dst + (size_t)dy * dw;
        for (int dx = 0; dx < dw; dx++) {
            const int sx = xofs[dx];
            const int sx1 = sx + 1 < sw ? sx + 1 : sx;
            const float a0 = alpha[2 * dx], a1 = alpha[2 * dx + 1];
            const float r0 = S0[sx] * a0 + S0[sx1] * a1; /* hresize */
            const float r1 = S1[sx] * a0 + S1[sx1] * a1;
            D[dx] = r0 * b0 + r1 * b1; /* vresize */
        }
    }
    free(xofs);
    free(alpha);
}

/* ---- cv::remap INTER_CUBIC, CV_32FC1, BORDER_CONSTANT(0) ------------------------------------- */
#define TAB 32
static float g_cubic1d[TAB][4];
static int g_tab_ready = 0;

static void init_cubic_tab(void) {
    if (g_tab_ready) return;
    const float A = -0.75f;
    for (int i = 0; i < TAB; i++) {
        const float x = (float)i * (1.f / TAB);
        float *c = g_cubic1d[i];
        c[0] = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A;
        c[1] = ((A + 2) * x - (A + 3)) * x * x + 1;
        c[2] = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1;
        c[3] = 1.f - c[0] - c[1] - c[2];
    }
    g_tab_ready = 1;
}

static void remap_cubic(const float *src, int h, int w, const float *mapx, const float *mapy, float *dst) {
    init_cubic_tab();
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++) {
        for (int x = 0; x < w; x++) {
            const size_t o = (size_t)y * w + x;
            const int ix = cv_round((double)(mapx[o] * (float)TAB));
            const int iy = cv_round((double)(mapy[o] * (float)TAB));
            const int sx = (ix >> 5) - 1, sy = (iy >> 5) - 1;
            const float *wx = g_cubic1d[ix & (TAB - 1)], *wy = g_cubic1d[iy & (TAB - 1)];
            float sum = 0.f;
            if (sx >= 0 && sy >= 0 && sx + 3 < w && sy + 3 < h) {
                const float *S = src + (size_t)sy * w + sx;
                for (int r = 0; r < 4; r++, S += w) {
                    const float w0 = wy[r] * wx[0], w1 = wy[r] * wx[1], w2 = wy[r] * wx[2], w3 = wy[r] * wx[3];
                    if (r == 0) sum = S[0] * w0 + S[1] * w1 + S[2] * w2 + S[3] * w3;
                    else sum += S[0] * w0 + S[1] * w1 + S[2] * w2 + S[3] * w3;
                }
            } else if (sx >= w || sx + 4 <= 0 || sy >= h || sy + 4 <= 0) {
                sum = 0.f; /* whole footprint outside -> borderValue */
            } else {
                for (int r = 0; r < 4; r++) {
                    const int yy = sy + r;
                    for (int c = 0; c < 4; c++) {
                        const int xx = sx + c;
                        const float v = (yy >= 0 && yy < h && xx >= 0 && xx < w) ? src[(size_t)yy * w + xx] : 0.f;
                        sum += v * (wy[r] * wx[c]);
                    }
                }
            }
            dst[o] = sum;
        }
    }
}

/* ---- in-tree stages (tvl1flow.cpp) ---------------------------------------------------------- */
static void centered_gradient(const float *s, int h, int w, float *dx, float *dy) { /* :718-770 */
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++) {
        const int ym = y > 0 ? y - 1 : 0, yp = y < h - 1 ? y + 1 : h - 1;
        for (int x = 0; x < w; x++) {
            const int xm = x > 0 ? x - 1 : 0, xp = x < w - 1 ? x + 1 : w - 1;
            dx[(size_t)y * w + x] = 0.5f * (s[(size_t)y * w + xp] - s[(size_t)y * w + xm]);
            dy[(size_t)y * w + x] = 0.5f * (s[(size_t)yp * w + x] - s[(size_t)ym * w + x]);
        }
    }
}

static void proc_one_scale(const tvl1_cpu_params *P, const float *I0, const float *I1, int h, int w, float *u1,
                           float *u2, float **ws) {
    const size_t n = (size_t)h * w;
    float *I1x = ws[0], *I1y = ws[1], *m1 = ws[2], *m2 = ws[3], *I1w = ws[4], *I1wx = ws[5], *I1wy = ws[6];
    float *grad = ws[7], *rho_c = ws[8], *v1 = ws[9], *v2 = ws[10];
    float *p11 = ws[11], *p12 = ws[12], *p21 = ws[13], *p22 = ws[14];
    float *div1 = ws[15], *div2 = ws[16], *u1x = ws[17], *u1y = ws[18], *u2x = ws[19], *u2y = ws[20];
    const float scaledEpsilon = (float)(P->epsilon * P->epsilon * (double)(h * w)); /* :1315 */
    centered_gradient(I1, h, w, I1x, I1y);
    memset(p11, 0, n * 4); memset(p12, 0, n * 4); memset(p21, 0, n * 4); memset(p22, 0, n * 4);
    const float l_t = (float)(P->lambda * P->theta), taut = (float)(P->tau / P->theta);
    const float theta = (float)P->theta;

    for (int wp = 0; wp < P->warps; ++wp) {
#pragma omp parallel for schedule(static)
        for (int y = 0; y < h; y++) /* buildFlowMap :651-668 */
            for (int x = 0; x < w; x++) {
                m1[(size_t)y * w + x] = x + u1[(size_t)y * w + x];
                m2[(size_t)y * w + x] = y + u2[(size_t)y * w + x];
            }
        remap_cubic(I1, h, w, m1, m2, I1w);  /* :1371-1373 */
        remap_cubic(I1x, h, w, m1, m2, I1wx);
        remap_cubic(I1y, h, w, m1, m2, I1wy);
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < n; i++) { /* calcGradRho :920-944 */
            const float Ix2 = I1wx[i] * I1wx[i], Iy2 = I1wy[i] * I1wy[i];
            grad[i] = Ix2 + Iy2;
            rho_c[i] = (I1w[i] - I1wx[i] * u1[i] - I1wy[i] * u2[i] - I0[i]);
        }
        float error = FLT_MAX;
        for (int no = 0; error > scaledEpsilon && no < P->outerIterations; ++no) {
            for (int ni = 0; error > scaledEpsilon && ni < P->innerIterations; ++ni) {
#pragma omp parallel for schedule(static)
                for (int y = 0; y < h; y++) {
                    for (int x = 0; x < w; x++) {
                        const size_t i = (size_t)y * w + x;
                        /* estimateV :992-1041 */
                        const float rho = rho_c[i] + (I1wx[i] * u1[i] + I1wy[i] * u2[i]);
                        float d1 = 0.f, d2 = 0.f;
                        if (rho < -l_t * grad[i]) { d1 = l_t * I1wx[i]; d2 = l_t * I1wy[i]; }
                        else if (rho > l_t * grad[i]) { d1 = -l_t * I1wx[i]; d2 = -l_t * I1wy[i]; }
                        else if (grad[i] > FLT_EPSILON) { const float fi = -rho / grad[i]; d1 = fi * I1wx[i]; d2 = fi * I1wy[i]; }
                        v1[i] = u1[i] + d1;
                        v2[i] = u2[i] + d2;
                        /* divergence :874-899 */
                        if (x > 0 && y > 0) {
                            div1[i] = (p11[i] - p11[i - 1]) + (p12[i] - p12[i - w]);
                            div2[i] = (p21[i] - p21[i - 1]) + (p22[i] - p22[i - w]);
                        } else if (y > 0) {
                            div1[i] = p11[i] + p12[i] - p12[i - w];
                            div2[i] = p21[i] + p22[i] - p22[i - w];
                        } else if (x > 0) {
                            div1[i] = p11[i] - p11[i - 1] + p12[i];
                            div2[i] = p21[i] - p21[i - 1] + p22[i];
                        } else {
                            div1[i] = p11[i] + p12[i];
                            div2[i] = p21[i] + p22[i];
                        }
                    }
                }
                /* estimateU :1074-1116 (the reference accumulates the error serially in float) */
                double err = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : err)
                for (int y = 0; y < h; y++) {
                    float e = 0.f;
                    for (int x = 0; x < w; x++) {
                        const size_t i = (size_t)y * w + x;
                        const float a = u1[i], b = u2[i];
                        u1[i] = v1[i] + theta * div1[i];
                        u2[i] = v2[i] + theta * div2[i];
                        e += (u1[i] - a) * (u1[i] - a) + (u2[i] - b) * (u2[i] - b);
                    }
                    err += e;
                }
                error = (float)err;
#pragma omp parallel for schedule(static)
                for (int y = 0; y < h; y++) { /* forwardGradient :804-840 */
                    for (int x = 0; x < w; x++) {
                        const size_t i = (size_t)y * w + x;
                        u1x[i] = x < w - 1 ? u1[i + 1] - u1[i] : 0.f;
                        u2x[i] = x < w - 1 ? u2[i + 1] - u2[i] : 0.f;
                        u1y[i] = y < h - 1 ? u1[i + w] - u1[i] : 0.f;
                        u2y[i] = y < h - 1 ? u2[i + w] - u2[i] : 0.f;
                    }
                }
#pragma omp parallel for schedule(static)
                for (size_t i = 0; i < n; i++) { /* estimateDualVariables :1140-1180 */
                    const float g1 = (float)hypot(u1x[i], u1y[i]);
                    const float g2 = (float)hypot(u2x[i], u2y[i]);
                    const float ng1 = 1.0f + taut * g1, ng2 = 1.0f + taut * g2;
                    p11[i] = (p11[i] + taut * u1x[i]) / ng1;
                    p12[i] = (p12[i] + taut * u1y[i]) / ng1;
                    p21[i] = (p21[i] + taut * u2x[i]) / ng2;
                    p22[i] = (p22[i] + taut * u2y[i]) / ng2;
                }
            }
        }
    }
}

/* ---- cv::medianBlur, CV_32FC1, ksize 3 or 5 -------------------------------------------------- */
/* OpenCV's float path is a sorting network over the k*k window with replicated borders, i.e. the exact
 * median; src must not alias dst (the glue passes copies). */
void tvl1_cpu_median_blur(const float *src, int h, int w, float *dst, int ksize) {
    const int r = ksize / 2;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++) {
        float win[25];
        for (int x = 0; x < w; x++) {
            int n = 0;
            for (int dy = -r; dy <= r; dy++) {
                int yy = y + dy; yy = yy < 0 ? 0 : (yy > h - 1 ? h - 1 : yy);
                for (int dx = -r; dx <= r; dx++) {
                    int xx = x + dx; xx = xx < 0 ? 0 : (xx > w - 1 ? w - 1 : xx);
                    win[n++] = src[(size_t)yy * w + xx];
                }
            }
            for (int i = 1; i < n; i++) { /* insertion sort: n <= 25 */
                const float v = win[i];
                int j = i - 1;
                while (j >= 0 && win[j] > v) { win[j + 1] = win[j]; j--; }
                win[j + 1] = v;
            }
            dst[(size_t)y * w + x] = win[n / 2];
        }
    }
}

/* debug / pinning entry points */
void tvl1_cpu_resize_linear_f(const float *src, int sh, int sw, float *dst, int dh, int dw, double f) {
    resize_linear(src, sh, sw, dst, dh, dw, f);
}
void tvl1_cpu_resize_linear(const float *src, int sh, int sw, float *dst, int dh, int dw) {
    resize_linear(src, sh, sw, dst, dh, dw, 0.0);
}
void tvl1_cpu_remap_cubic(const float *src, int h, int w, const float *mapx, const float *mapy, float *dst) {
    remap_cubic(src, h, w, mapx, mapy, dst);
}

/* OpticalFlowDual_TVL1::calc (tvl1flow.cpp:402-533), gamma = 0, medianFiltering = 1.
 * I0/I1: float32 already scaled to 0..255 (the caller applies the x1 / x255 rule of :429-430).
 * flow: interleaved (u, v) float32.  Returns 0 on success. */
/* thread count for every parallel region (bench.py passes the CPUs this process may actually use) */
int tvl1_cpu_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

int tvl1_cpu_calc(const tvl1_cpu_params *P, const float *I0, const float *I1, int rows, int cols, float *flow) {
    if (!P || P->nscales <= 0 || P->gamma != 0.0 || P->medianFiltering > 1 || P->useInitialFlow) return -1;
    int nscales = P->nscales;
    int *hs = (int *)malloc(sizeof(int) * nscales), *wsz = (int *)malloc(sizeof(int) * nscales);
    float **I0s = (float **)calloc(nscales, sizeof(float *)), **I1s = (float **)calloc(nscales, sizeof(float *));
    float **u1s = (float **)calloc(nscales, sizeof(float *)), **u2s = (float **)calloc(nscales, sizeof(float *));
    const size_t n0 = (size_t)rows * cols;
    hs[0] = rows; wsz[0] = cols;
    I0s[0] = alloc_rows(rows, cols, 0); I1s[0] = alloc_rows(rows, cols, 0);
    copy_rows(I0s[0], I0, rows, cols); copy_rows(I1s[0], I1, rows, cols);
    u1s[0] = alloc_rows(rows, cols, 1); u2s[0] = alloc_rows(rows, cols, 1);
    int built = 1;
    for (int s = 1; s < nscales; ++s) { /* :477-503 */
        hs[s] = cv_round(hs[s - 1] * P->scaleStep);
        wsz[s] = cv_round(wsz[s - 1] * P->scaleStep);
        I0s[s] = alloc_rows(hs[s], wsz[s], 0); I1s[s] = alloc_rows(hs[s], wsz[s], 0);
        resize_linear(I0s[s - 1], hs[s - 1], wsz[s - 1], I0s[s], hs[s], wsz[s], P->scaleStep);
        resize_linear(I1s[s - 1], hs[s - 1], wsz[s - 1], I1s[s], hs[s], wsz[s], P->scaleStep);
        built = s + 1;
        if (wsz[s] < 16 || hs[s] < 16) { nscales = s; break; }
        u1s[s] = alloc_rows(hs[s], wsz[s], 1); u2s[s] = alloc_rows(hs[s], wsz[s], 1);
    }
    float *ws[21];
    for (int i = 0; i < 21; i++) ws[i] = alloc_rows(rows, cols, 0);
    for (int s = nscales - 1; s >= 0; --s) { /* :510-529 */
        proc_one_scale(P, I0s[s], I1s[s], hs[s], wsz[s], u1s[s], u2s[s], ws);
        if (s == 0) break;
        resize_linear(u1s[s], hs[s], wsz[s], u1s[s - 1], hs[s - 1], wsz[s - 1], 0.0);
        resize_linear(u2s[s], hs[s], wsz[s], u2s[s - 1], hs[s - 1], wsz[s - 1], 0.0);
        const float inv = (float)(1 / P->scaleStep);
        const size_t n = (size_t)hs[s - 1] * wsz[s - 1];
        for (size_t i = 0; i < n; i++) { u1s[s - 1][i] *= inv; u2s[s - 1][i] *= inv; }
    }
    for (size_t i = 0; i < n0; i++) { flow[2 * i] = u1s[0][i]; flow[2 * i + 1] = u2s[0][i]; }
    for (int i = 0; i < 21; i++) free(ws[i]);
    for (int s = 0; s < built; ++s) { free(I0s[s]); free(I1s[s]); free(u1s[s]); free(u2s[s]); }
    free(I0s); free(I1s); free(u1s); free(u2s); free(hs); free(wsz);
    return 0;
}
