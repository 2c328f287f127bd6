// denselk.cu -- cv::cuda::DensePyrLKOpticalFlow re-implemented for sm_100a.
//
// Reference being replaced (paths relative to /root/reference/modules/cudaoptflow):
//   host  src/pyrlk.cpp:238-299 (PyrLKOpticalFlowBase::dense), :379-392 (calc)
//   dev   src/cuda/pyrlk.cu:709-855 (denseKernel), :535-554 (TextureLinear)
//   pyramid: cudawarping/src/cuda/pyr_down.cu:55-173
//
// Semantics kept, because no upstream test decides otherwise:
//  * the I patch and its Scharr derivatives are truncated to int, A = sum dI dI^T and b accumulate in
//    32-bit integers (two's-complement wrap-around, which is what the reference's GPU code does);
//  * J is sampled with bilinear filtering and truncated to int; the bilinear weights are quantised to
//    8 fractional bits like the texture unit the reference samples through (CUDA programming guide,
//    "linear filtering": 9-bit fixed point with 8 bits of fractional value);
//  * a pixel whose matrix is singular or whose track leaves the image is NOT written: it keeps what
//    the ping-pong buffer held (zeros, or the value written two levels earlier at the same index).
// Differences by design: window sizes / iteration count are kernel parameters (the reference uploads
// them into global __constant__ symbols, pyrlk.cu:857-869, racing between instances); the pyramid and
// the ping-pong flow buffers live in the handle's arena; one launch per level on the caller's stream.
//
// Kernel classes: 0 lk_dense (compute bound: win^2 x iters bilinear taps per pixel), 1 pyramid, 2 misc.
#include "common.cuh"

#include <cstring>
#include <new>
#include <vector>

namespace b2f {

namespace {

enum { CLS_LK = 0, CLS_PYR = 1, CLS_MISC = 2 };
constexpr int LK_BX = 16, LK_BY = 16;  // pyrlk.cu:921

// point / clamp fetch of the float image at texel centre (texI(y, x) with x = ix + 0.5)
__device__ __forceinline__ float lk_texel(const Plane &P, int rows, int cols, int y, int x) {
    return __ldg(&P.at(clampi(y, 0, rows - 1), clampi(x, 0, cols - 1)));
}

// bilinear blend with every operation rounded separately (no FMA contraction), left to right: the order the
// numpy model (oracle/denselk_model.py) evaluates, so generic kernel, fast kernel and model agree bit for bit
__device__ __forceinline__ float lk_blend(float ax, float ay, float t00, float t01, float t10, float t11) {
    const float bx = __fsub_rn(1.f, ax), by = __fsub_rn(1.f, ay);
    float r = __fmul_rn(__fmul_rn(bx, by), t00);
    r = __fadd_rn(r, __fmul_rn(__fmul_rn(ax, by), t01));
    r = __fadd_rn(r, __fmul_rn(__fmul_rn(bx, ay), t10));
    r = __fadd_rn(r, __fmul_rn(__fmul_rn(ax, ay), t11));
    return r;
}

// texture-unit model of one axis: texel index of the lower tap (clamped copies for both taps) and the weight
// of the upper tap with 8 fractional bits
__device__ __forceinline__ void lk_axis(float coord, int n, int &ia, int &ic, float &a) {
    const float b = __fsub_rn(coord, 0.5f);
    const float f = floorf(b);
    a = __fmul_rn(floorf(__fadd_rn(__fmul_rn(__fsub_rn(b, f), 256.f), 0.5f)), 1.f / 256.f);
    const int i0 = (int)fminf(fmaxf(f, -2.f), (float)n + 1.f);
    ia = clampi(i0, 0, n - 1);
    ic = clampi(i0 + 1, 0, n - 1);
}

// unnormalised, linear, clamp-addressed fetch at (x, y) in texture coordinates (texel centres at +0.5)
__device__ __forceinline__ float lk_bilinear(const Plane &P, int rows, int cols, float y, float x) {
    int xa, xc, ya, yc;
    float ax, ay;
    lk_axis(x, cols, xa, xc, ax);
    lk_axis(y, rows, ya, yc, ay);
    return lk_blend(ax, ay, __ldg(&P.at(ya, xa)), __ldg(&P.at(ya, xc)), __ldg(&P.at(yc, xa)), __ldg(&P.at(yc, xc)));
}

__global__ void __launch_bounds__(LK_BX *LK_BY) k_lk_dense(Plane I, Plane J, Plane u, Plane v, Plane prevU, Plane prevV,
                                                           int rows, int cols, int win_x, int win_y, int half_x,
                                                           int half_y, int iters) {
    extern __shared__ int lk_smem[];
    const int patchW = LK_BX + 2 * half_x, patchH = LK_BY + 2 * half_y;
    int *I_patch = lk_smem;
    int *dIdx_patch = I_patch + patchW * patchH;
    int *dIdy_patch = dIdx_patch + patchW * patchH;
    const int xBase = blockIdx.x * LK_BX, yBase = blockIdx.y * LK_BY;

    for (int i = threadIdx.y; i < patchH; i += LK_BY) {
        for (int j = threadIdx.x; j < patchW; j += LK_BX) {
            const int px = xBase - half_x + j, py = yBase - half_y + i;  // texel indices (texture coord - 0.5)
            auto T = [&](int dy, int dx) { return lk_texel(I, rows, cols, py + dy, px + dx); };
            I_patch[i * patchW + j] = (int)T(0, 0);
            // Scharr derivative, evaluated in float then truncated (pyrlk.cu:733-740)
            dIdx_patch[i * patchW + j] =
                (int)(3 * T(-1, 1) + 10 * T(0, 1) + 3 * T(1, 1) - (3 * T(-1, -1) + 10 * T(0, -1) + 3 * T(1, -1)));
            dIdy_patch[i * patchW + j] =
                (int)(3 * T(1, -1) + 10 * T(1, 0) + 3 * T(1, 1) - (3 * T(-1, -1) + 10 * T(-1, 0) + 3 * T(-1, 1)));
        }
    }
    __syncthreads();

    const int x = xBase + threadIdx.x, y = yBase + threadIdx.y;
    if (x >= cols || y >= rows) return;

    unsigned A11i = 0, A12i = 0, A22i = 0;  // int32 accumulation with wrap-around
    for (int i = 0; i < win_y; ++i) {
        for (int j = 0; j < win_x; ++j) {
            const int dIdx = dIdx_patch[(threadIdx.y + i) * patchW + (threadIdx.x + j)];
            const int dIdy = dIdy_patch[(threadIdx.y + i) * patchW + (threadIdx.x + j)];
            A11i += (unsigned)(dIdx * dIdx);
            A12i += (unsigned)(dIdx * dIdy);
            A22i += (unsigned)(dIdy * dIdy);
        }
    }
    float A11 = (float)(int)A11i, A12 = (float)(int)A12i, A22 = (float)(int)A22i;
    float D = A11 * A22 - A12 * A12;
    if (D < FLT_EPSILON) return;  // singular: keep the buffer's previous content (pyrlk.cu:775-780)
    D = 1.f / D;
    A11 *= D;
    A12 *= D;
    A22 *= D;

    float nx = x + prevU.at(y / 2, x / 2) * 2.0f;
    float ny = y + prevV.at(y / 2, x / 2) * 2.0f;
    for (int k = 0; k < iters; ++k) {
        if (nx < 0 || nx >= cols || ny < 0 || ny >= rows) return;  // left the image: no write (pyrlk.cu:793-798)
        unsigned b1 = 0, b2 = 0;
        for (int i = 0; i < win_y; ++i) {
            const float jy = __fadd_rn(__fadd_rn(__fsub_rn(ny, (float)half_y), (float)i), 0.5f);
            for (int j = 0; j < win_x; ++j) {
                const int Iv = I_patch[(threadIdx.y + i) * patchW + threadIdx.x + j];
                const int Jv = (int)lk_bilinear(J, rows, cols, jy, __fadd_rn(__fadd_rn(__fsub_rn(nx, (float)half_x), (float)j), 0.5f));
                const int diff = (Jv - Iv) * 32;
                const int dIdx = dIdx_patch[(threadIdx.y + i) * patchW + (threadIdx.x + j)];
                const int dIdy = dIdy_patch[(threadIdx.y + i) * patchW + (threadIdx.x + j)];
                b1 += (unsigned)(diff * dIdx);
                b2 += (unsigned)(diff * dIdy);
            }
        }
        const float fb1 = (float)(int)b1, fb2 = (float)(int)b2;
        const float dx = A12 * fb2 - A22 * fb1;
        const float dy = A12 * fb1 - A11 * fb2;
        nx += dx;
        ny += dy;
        if (fabsf(dx) < 0.01f && fabsf(dy) < 0.01f) break;
    }
    u.at(y, x) = nx - x;
    v.at(y, x) = ny - y;
}

// ---------------------------------------------------------------------------------------------
// Fast path for a compile-time window width (the default 13): the x half of every bilinear fetch (tap
// indices + weight) is the same for all window rows, so it is computed once per iteration into registers;
// the y half once per window row.  Neighbouring taps share texels (the 13x13 window touches 14x14 texels),
// so a window row costs 14 loads -- the upper texel row is kept in registers for the next window row --
// instead of 52: the per-tap kernel is bound by L1 wavefronts (676 loads per pixel per iteration).
// dI/dx and dI/dy are packed into one 32-bit word (|Scharr of an 8-bit image| <= 4080 fits 16 bits).
// Same arithmetic per tap as k_lk_dense -> bit-identical output.
// ---------------------------------------------------------------------------------------------
template <int WX>
__global__ void __launch_bounds__(LK_BX *LK_BY) k_lk_dense_fast(Plane I, Plane J, Plane u, Plane v, Plane prevU, Plane prevV,
                                                                int rows, int cols, int win_y, int half_y, int iters) {
    extern __shared__ int lk_smem[];
    constexpr int half_x = (WX - 1) / 2;
    constexpr int patchW = LK_BX + 2 * half_x;
    const int patchH = LK_BY + 2 * half_y;
    int *I_patch = lk_smem;
    int *D_patch = I_patch + patchW * patchH;
    const int xBase = blockIdx.x * LK_BX, yBase = blockIdx.y * LK_BY;

    for (int i = threadIdx.y; i < patchH; i += LK_BY) {
        for (int j = threadIdx.x; j < patchW; j += LK_BX) {
            const int px = xBase - half_x + j, py = yBase - half_y + i;
            auto T = [&](int dy, int dx) { return lk_texel(I, rows, cols, py + dy, px + dx); };
            I_patch[i * patchW + j] = (int)T(0, 0);
            const int dx = (int)(3 * T(-1, 1) + 10 * T(0, 1) + 3 * T(1, 1) - (3 * T(-1, -1) + 10 * T(0, -1) + 3 * T(1, -1)));
            const int dy = (int)(3 * T(1, -1) + 10 * T(1, 0) + 3 * T(1, 1) - (3 * T(-1, -1) + 10 * T(-1, 0) + 3 * T(-1, 1)));
            D_patch[i * patchW + j] = (dx & 0xffff) | (dy << 16);
        }
    }
    __syncthreads();

    const int x = xBase + threadIdx.x, y = yBase + threadIdx.y;
    if (x >= cols || y >= rows) return;
    const int *Ip = I_patch + threadIdx.y * patchW + threadIdx.x;
    const int *Dp = D_patch + threadIdx.y * patchW + threadIdx.x;

    unsigned A11i = 0, A12i = 0, A22i = 0;
    for (int i = 0; i < win_y; ++i) {
#pragma unroll
        for (int j = 0; j < WX; ++j) {
            const int d = Dp[i * patchW + j];
            const int dIdx = (int)(short)(d & 0xffff), dIdy = d >> 16;
            A11i += (unsigned)(dIdx * dIdx);
            A12i += (unsigned)(dIdx * dIdy);
            A22i += (unsigned)(dIdy * dIdy);
        }
    }
    float A11 = (float)(int)A11i, A12 = (float)(int)A12i, A22 = (float)(int)A22i;
    float D = A11 * A22 - A12 * A12;
    if (D < FLT_EPSILON) return;
    D = 1.f / D;
    A11 *= D;
    A12 *= D;
    A22 *= D;

    float nx = x + prevU.at(y / 2, x / 2) * 2.0f;
    float ny = y + prevV.at(y / 2, x / 2) * 2.0f;
    for (int k = 0; k < iters; ++k) {
        if (nx < 0 || nx >= cols || ny < 0 || ny >= rows) return;
        // x half of every fetch, once per iteration.  Adjacent taps share a texel column (the upper tap of
        // tap j is the lower tap of tap j+1) unless float rounding of the tap coordinate breaks the chain;
        // then the iteration falls back to per-tap fetches.
        int col[WX + 1];
        float ax[WX];
        bool chained = true;
        {
            int prev = 0;
#pragma unroll
            for (int j = 0; j < WX; ++j) {
                int xa, xc;
                lk_axis(__fadd_rn(__fadd_rn(__fsub_rn(nx, (float)half_x), (float)j), 0.5f), cols, xa, xc, ax[j]);
                if (j > 0) chained = chained && (xa == prev);
                col[j] = xa;
                prev = xc;
            }
            col[WX] = prev;
        }
        unsigned b1 = 0, b2 = 0;
        if (chained) {
            float lo[WX + 1], hi[WX + 1];
            int prev_yc = -1;
            for (int i = 0; i < win_y; ++i) {
                int ya, yc;
                float ay;
                lk_axis(__fadd_rn(__fadd_rn(__fsub_rn(ny, (float)half_y), (float)i), 0.5f), rows, ya, yc, ay);
                const float *ra = J.row(ya), *rc = J.row(yc);
                if (i > 0 && ya == prev_yc) {  // the usual case: this row's lower texel row is the last upper one
#pragma unroll
                    for (int j = 0; j <= WX; ++j) lo[j] = hi[j];
                } else {
#pragma unroll
                    for (int j = 0; j <= WX; ++j) lo[j] = __ldg(ra + col[j]);
                }
#pragma unroll
                for (int j = 0; j <= WX; ++j) hi[j] = __ldg(rc + col[j]);
                prev_yc = yc;
#pragma unroll
                for (int j = 0; j < WX; ++j) {
                    const int Jv = (int)lk_blend(ax[j], ay, lo[j], lo[j + 1], hi[j], hi[j + 1]);
                    const int diff = (Jv - Ip[i * patchW + j]) * 32;
                    const int d = Dp[i * patchW + j];
                    b1 += (unsigned)(diff * (int)(short)(d & 0xffff));
                    b2 += (unsigned)(diff * (d >> 16));
                }
            }
        } else {
            for (int i = 0; i < win_y; ++i) {
                const float jy = __fadd_rn(__fadd_rn(__fsub_rn(ny, (float)half_y), (float)i), 0.5f);
                for (int j = 0; j < WX; ++j) {
                    const int Jv = (int)lk_bilinear(J, rows, cols, jy, __fadd_rn(__fadd_rn(__fsub_rn(nx, (float)half_x), (float)j), 0.5f));
                    const int diff = (Jv - Ip[i * patchW + j]) * 32;
                    const int d = Dp[i * patchW + j];
                    b1 += (unsigned)(diff * (int)(short)(d & 0xffff));
                    b2 += (unsigned)(diff * (d >> 16));
                }
            }
        }
        const float fb1 = (float)(int)b1, fb2 = (float)(int)b2;
        const float dx = A12 * fb2 - A22 * fb1;
        const float dy = A12 * fb1 - A11 * fb2;
        nx += dx;
        ny += dy;
        if (fabsf(dx) < 0.01f && fabsf(dy) < 0.01f) break;
    }
    u.at(y, x) = nx - x;
    v.at(y, x) = ny - y;
}

// ---------------------------------------------------------------------------------------------
// host engine
// ---------------------------------------------------------------------------------------------
struct LLevel {
    int rows = 0, cols = 0;
    Plane I, J;
};

class DenseLKEngine : public b2f_handle {
public:
    explicit DenseLKEngine(const b2f_denselk_params &p) : P(p) { algo = ALGO_DENSELK; }
    b2f_denselk_params P;

    int calc(const b2f_image *I0, const b2f_image *I1, b2f_image *flow, cudaStream_t s) override;
    int set_param(int id, double v) override;
    int get_param(int id, double *v) const override;
    const char *default_name() const override { return "DenseOpticalFlow.DensePyrLKOpticalFlow"; }
    const char *class_name(int cls) const override {
        static const char *n[] = {"lk_dense", "lk_pyramid", "lk_misc"};
        return (cls >= 0 && cls < 3) ? n[cls] : "";
    }
    size_t workspace_bytes(int rows, int cols, int type) override {
        (void)type;
        Layout L;
        return layout(rows, cols, true, L);
    }
    // rejected pixels are never written (pyrlk.cu:709-855): what the caller's flow held stays visible
    bool reads_flow() const override { return true; }

private:
    struct Layout {
        int rows = 0, cols = 0, max_level = -1;
        std::vector<LLevel> levels;
        Plane uP[2], vP[2];  // full-res ping-pong buffers; coarse levels use their top-left corner (pyrlk.cpp:261-275)
    };
    Layout L_;
    size_t layout(int rows, int cols, bool counting, Layout &L);
};

size_t DenseLKEngine::layout(int rows, int cols, bool counting, Layout &L) {
    Arena tmp;
    Arena &A = counting ? tmp : arena;
    A.begin(counting);
    L.rows = rows;
    L.cols = cols;
    L.max_level = P.max_level;
    L.levels.clear();
    int r = rows, c = cols;
    for (int l = 0; l <= P.max_level; ++l) {
        if (l > 0) {  // pyrDown: ((rows+1)/2, (cols+1)/2), pyramids.cpp:66-93
            r = (r + 1) / 2;
            c = (c + 1) / 2;
        }
        LLevel lv;
        lv.rows = r;
        lv.cols = c;
        lv.I = A.plane(r, c);
        lv.J = A.plane(r, c);
        L.levels.push_back(lv);
    }
    for (int i = 0; i < 2; ++i) {
        L.uP[i] = A.plane(rows, cols);
        L.vP[i] = A.plane(rows, cols);
    }
    return A.used();
}

int DenseLKEngine::calc(const b2f_image *I0, const b2f_image *I1, b2f_image *flow, cudaStream_t s) {
    // preconditions, pyrlk.cpp:240-243
    if (I0->type != B2F_8UC1 || I1->type != B2F_8UC1) return B2F_UNSUPPORTED_TYPE;
    if (I0->rows != I1->rows || I0->cols != I1->cols) return B2F_SIZE_MISMATCH;
    if (!flow_type_ok(flow)) return B2F_UNSUPPORTED_TYPE;
    if (flow->rows != I0->rows || flow->cols != I0->cols) return B2F_SIZE_MISMATCH;
    if (P.max_level < 0 || !(P.win_width > 2 && P.win_height > 2) || P.iters < 0) return B2F_BAD_ARG;
    if (I0->step < (size_t)I0->cols || I1->step < (size_t)I1->cols || !flow_step_ok(flow)) return B2F_BAD_ARG;
    const int rows = I0->rows, cols = I0->cols;
    Ctx c = make_ctx(s);
    if (!(L_.rows == rows && L_.cols == cols && L_.max_level == P.max_level && arena.capacity() > 0)) {
        Layout tmp;
        const size_t need = layout(rows, cols, true, tmp);
        c.check(arena.reserve(need));
        if (c.ok()) layout(rows, cols, false, L_);
    }
    if (!c.ok()) return finish(c, s);
    stats.levels = P.max_level + 1;
    stats.iterations_run = 0;

    const int half_x = (P.win_width - 1) / 2, half_y = (P.win_height - 1) / 2;  // pyrlk.cpp:110-111
    const size_t smem = sizeof(int) * 3 * (LK_BX + 2 * half_x) * (LK_BY + 2 * half_y);
    if (smem > 200 * 1024) return B2F_BAD_ARG;
    if (smem > 48 * 1024) {
        c.check(cudaFuncSetAttribute(k_lk_dense, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        c.check(cudaFuncSetAttribute(k_lk_dense_fast<13>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }

    const ImageView v0{I0->data, I0->step, rows, cols, I0->type};
    const ImageView v1{I1->data, I1->step, rows, cols, I1->type};
    const ImageView vf = flow_view(flow, rows, cols);
    convert_pair(c, CLS_PYR, v0, v1, L_.levels[0].I, L_.levels[0].J, 1.0f);  // convertTo(CV_32F), pyrlk.cpp:252-253
    for (int l = 1; l <= P.max_level; ++l) {
        const LLevel &a = L_.levels[l - 1], &b = L_.levels[l];
        pyr_down(c, CLS_PYR, a.I, a.rows, a.cols, b.I, b.rows, b.cols);
        pyr_down(c, CLS_PYR, a.J, a.rows, a.cols, b.J, b.rows, b.cols);
    }
    for (int i = 0; i < 2; ++i) {  // pyrlk.cpp:272-275
        fill_plane(c, L_.uP[i], rows, cols, 0.f);
        fill_plane(c, L_.vP[i], rows, cols, 0.f);
    }
    int idx = 0;
    for (int l = P.max_level; l >= 0; --l) {
        const int idx2 = (idx + 1) & 1;
        const LLevel &lv = L_.levels[l];
        const dim3 grid(div_up(lv.cols, LK_BX), div_up(lv.rows, LK_BY));
        if (P.win_width == 13 && knobs.kernel_path != 1) {
            const size_t smem2 = sizeof(int) * 2 * (LK_BX + 2 * half_x) * (LK_BY + 2 * half_y);
            B2F_LAUNCH(c, CLS_LK, 24.0 * lv.rows * lv.cols, k_lk_dense_fast<13>, grid, dim3(LK_BX, LK_BY), smem2, lv.I, lv.J,
                       L_.uP[idx], L_.vP[idx], L_.uP[idx2], L_.vP[idx2], lv.rows, lv.cols, P.win_height, half_y, P.iters);
        } else {
            B2F_LAUNCH(c, CLS_LK, 24.0 * lv.rows * lv.cols, k_lk_dense, grid, dim3(LK_BX, LK_BY), smem, lv.I, lv.J, L_.uP[idx],
                       L_.vP[idx], L_.uP[idx2], L_.vP[idx2], lv.rows, lv.cols, P.win_width, P.win_height, half_x, half_y,
                       P.iters);
        }
        if (l > 0) idx = idx2;
    }
    merge_flow(c, CLS_MISC, L_.uP[idx], L_.vP[idx], vf);  // pyrlk.cpp:297-298,390-391
    return finish(c, s);
}

int DenseLKEngine::set_param(int id, double v) {
    switch (id) {
        case B2F_LK_WIN_WIDTH: P.win_width = static_cast<int>(v); break;
        case B2F_LK_WIN_HEIGHT: P.win_height = static_cast<int>(v); break;
        case B2F_LK_MAX_LEVEL: P.max_level = static_cast<int>(v); break;
        case B2F_LK_ITERS: P.iters = static_cast<int>(v); break;
        case B2F_LK_USE_INITIAL_FLOW: P.use_initial_flow = v != 0; break;  // stored, unused by the dense path (as upstream)
        default: return B2F_BAD_ARG;
    }
    return B2F_OK;
}

int DenseLKEngine::get_param(int id, double *v) const {
    switch (id) {
        case B2F_LK_WIN_WIDTH: *v = P.win_width; break;
        case B2F_LK_WIN_HEIGHT: *v = P.win_height; break;
        case B2F_LK_MAX_LEVEL: *v = P.max_level; break;
        case B2F_LK_ITERS: *v = P.iters; break;
        case B2F_LK_USE_INITIAL_FLOW: *v = P.use_initial_flow; break;
        default: return B2F_BAD_ARG;
    }
    return B2F_OK;
}

}  // namespace

}  // namespace b2f

extern "C" {

void b2f_denselk_default_params(b2f_denselk_params *p) {
    if (!p) return;
    p->win_width = 13;
    p->win_height = 13;
    p->max_level = 3;
    p->iters = 30;
    p->use_initial_flow = 0;
}

int b2f_denselk_create(const b2f_denselk_params *p, b2f_handle **out) {
    if (!out) return B2F_BAD_ARG;
    b2f_denselk_params d;
    b2f_denselk_default_params(&d);
    if (p) d = *p;
    *out = new (std::nothrow) b2f::DenseLKEngine(d);
    return *out ? B2F_OK : B2F_OUT_OF_MEMORY;
}

}  // extern "C"
