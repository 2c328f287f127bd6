// tvl1_blocked.cu -- see tvl1_blocked.cuh for the design.
#include "tvl1_blocked.cuh"

namespace b2f {

namespace {

constexpr int R = TVL1_REGION;            // region edge (pixels)
constexpr int NT = TVL1_BLOCK_THREADS;    // 512
constexpr int PLANE_F = R * R;            // floats per staged plane
constexpr int N_IN = 10;                  // Ix, Iy, grad, rho_c, u1, u2, p11, p12, p21, p22
constexpr int N_OUT = 6;                  // u1, u2, p11, p12, p21, p22
constexpr int EX_F = 32 * R;              // floats per exchange array (32 thread-rows x 64)
constexpr size_t SMEM_BYTES = sizeof(float) * (size_t)(N_IN * PLANE_F + 4 * EX_F);

struct Regs {
    float Ix[2][4], Iy[2][4], gr[2][4], rc[2][4];
    float u1[2][4], u2[2][4], p11[2][4], p12[2][4], p21[2][4], p22[2][4];
};

__device__ __forceinline__ void ld4(const float *s, float (&d)[4]) {
    const float4 v = *reinterpret_cast<const float4 *>(s);
    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
}
__device__ __forceinline__ void st4(float *s, const float (&d)[4]) {
    *reinterpret_cast<float4 *>(s) = make_float4(d[0], d[1], d[2], d[3]);
}

// K iterations on the register-resident region.  ex = 4 exchange arrays [u1 | u2 | p12 | p22].
// gxb/gyb: global coordinates of the thread's first pixel; W/H image size (BORDER only).
template <bool BORDER>
__device__ __forceinline__ void tile_iterate(Regs &r, float *ex, int iters, const Tvl1Scalars k, int lx, int tr,
                                             int gxb, int gyb, int W, int H) {
    float *ex_u1 = ex, *ex_u2 = ex + EX_F, *ex_p12 = ex + 2 * EX_F, *ex_p22 = ex + 3 * EX_F;
    const int mine = tr * R + 4 * lx;
    const int up = max(tr - 1, 0) * R + 4 * lx;
    const int dn = min(tr + 1, 31) * R + 4 * lx;

    // publish the bottom rows of p12/p22 for the first primal update
    st4(ex_p12 + mine, r.p12[1]);
    st4(ex_p22 + mine, r.p22[1]);
    __syncthreads();

    for (int it = 0; it < iters; ++it) {
        // ---------------- primal update (estimateU) ----------------
        float up12[4], up22[4];
        ld4(ex_p12 + up, up12);
        ld4(ex_p22 + up, up22);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const float l11 = __shfl_up_sync(0xffffffffu, r.p11[j][3], 1, 16);
            const float l21 = __shfl_up_sync(0xffffffffu, r.p21[j][3], 1, 16);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float pl11 = i ? r.p11[j][i ? i - 1 : 0] : l11;
                float pl21 = i ? r.p21[j][i ? i - 1 : 0] : l21;
                float pu12 = j ? r.p12[0][i] : up12[i];
                float pu22 = j ? r.p22[0][i] : up22[i];
                if (BORDER) {
                    if (gxb + i == 0) { pl11 = 0.f; pl21 = 0.f; }
                    if (gyb + j == 0) { pu12 = 0.f; pu22 = 0.f; }
                }
                float a, b;
                tvl1_update_u(k, r.Ix[j][i], r.Iy[j][i], r.gr[j][i], r.rc[j][i], r.u1[j][i], r.u2[j][i],
                              r.p11[j][i], pl11, r.p12[j][i], pu12, r.p21[j][i], pl21, r.p22[j][i], pu22, a, b);
                r.u1[j][i] = a;
                r.u2[j][i] = b;
            }
        }
        st4(ex_u1 + mine, r.u1[0]);
        st4(ex_u2 + mine, r.u2[0]);
        __syncthreads();

        // ---------------- dual update (estimateDualVariables) ----------------
        float dn1[4], dn2[4];
        ld4(ex_u1 + dn, dn1);
        ld4(ex_u2 + dn, dn2);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const float r1 = __shfl_down_sync(0xffffffffu, r.u1[j][0], 1, 16);
            const float r2 = __shfl_down_sync(0xffffffffu, r.u2[j][0], 1, 16);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float c1 = r.u1[j][i], c2 = r.u2[j][i];
                const float ur1 = i < 3 ? r.u1[j][i < 3 ? i + 1 : 3] : r1;
                const float ur2 = i < 3 ? r.u2[j][i < 3 ? i + 1 : 3] : r2;
                const float ud1 = j == 0 ? r.u1[1][i] : dn1[i];
                const float ud2 = j == 0 ? r.u2[1][i] : dn2[i];
                float ux1 = __fsub_rn(ur1, c1), uy1 = __fsub_rn(ud1, c1);
                float ux2 = __fsub_rn(ur2, c2), uy2 = __fsub_rn(ud2, c2);
                if (BORDER) {
                    if (gxb + i == W - 1) { ux1 = 0.f; ux2 = 0.f; }
                    if (gyb + j == H - 1) { uy1 = 0.f; uy2 = 0.f; }
                }
                tvl1_update_p(k.taut, ux1, uy1, r.p11[j][i], r.p12[j][i]);
                tvl1_update_p(k.taut, ux2, uy2, r.p21[j][i], r.p22[j][i]);
            }
        }
        st4(ex_p12 + mine, r.p12[1]);
        st4(ex_p22 + mine, r.p22[1]);
        __syncthreads();
    }
}

struct InPlanes {
    Plane p[N_IN];
};
struct OutPlanes {
    Plane p[N_OUT];
};

// One 64x64 region per CTA; plain coalesced loads/stores through the shared staging buffer.
__global__ void __launch_bounds__(NT, 1)
    k_tvl1_blocked(InPlanes in, OutPlanes out, int rows, int cols, Tvl1Scalars k, int iters, int tile) {
    extern __shared__ __align__(16) float smem[];
    float *stage = smem;
    float *ex = smem + N_IN * PLANE_F;

    const int tid = threadIdx.x;
    const int lx = tid & 15, tr = tid >> 4;
    const int gx0 = blockIdx.x * tile - iters;  // region origin (may be negative)
    const int gy0 = blockIdx.y * tile - iters;

    // ---- stage the region (zero fill outside the image) ----
#pragma unroll 1
    for (int pl = 0; pl < N_IN; ++pl) {
        const Plane P = in.p[pl];
        float *dst = stage + pl * PLANE_F;
#pragma unroll
        for (int n = 0; n < PLANE_F / NT; ++n) {
            const int idx = tid + n * NT;
            const int ry = idx >> 6, rx = idx & 63;
            const int gy = gy0 + ry, gx = gx0 + rx;
            float v = 0.f;
            if (gx >= 0 && gy >= 0 && gx < cols && gy < rows) v = __ldg(&P.at(gy, gx));
            dst[idx] = v;
        }
    }
    __syncthreads();

    Regs r;
    {
        const int o0 = (2 * tr) * R + 4 * lx, o1 = o0 + R;
        ld4(stage + 0 * PLANE_F + o0, r.Ix[0]);  ld4(stage + 0 * PLANE_F + o1, r.Ix[1]);
        ld4(stage + 1 * PLANE_F + o0, r.Iy[0]);  ld4(stage + 1 * PLANE_F + o1, r.Iy[1]);
        ld4(stage + 2 * PLANE_F + o0, r.gr[0]);  ld4(stage + 2 * PLANE_F + o1, r.gr[1]);
        ld4(stage + 3 * PLANE_F + o0, r.rc[0]);  ld4(stage + 3 * PLANE_F + o1, r.rc[1]);
        ld4(stage + 4 * PLANE_F + o0, r.u1[0]);  ld4(stage + 4 * PLANE_F + o1, r.u1[1]);
        ld4(stage + 5 * PLANE_F + o0, r.u2[0]);  ld4(stage + 5 * PLANE_F + o1, r.u2[1]);
        ld4(stage + 6 * PLANE_F + o0, r.p11[0]); ld4(stage + 6 * PLANE_F + o1, r.p11[1]);
        ld4(stage + 7 * PLANE_F + o0, r.p12[0]); ld4(stage + 7 * PLANE_F + o1, r.p12[1]);
        ld4(stage + 8 * PLANE_F + o0, r.p21[0]); ld4(stage + 8 * PLANE_F + o1, r.p21[1]);
        ld4(stage + 9 * PLANE_F + o0, r.p22[0]); ld4(stage + 9 * PLANE_F + o1, r.p22[1]);
    }

    const int gxb = gx0 + 4 * lx, gyb = gy0 + 2 * tr;
    const bool border = gx0 <= 0 || gy0 <= 0 || gx0 + R >= cols || gy0 + R >= rows;
    if (border)
        tile_iterate<true>(r, ex, iters, k, lx, tr, gxb, gyb, cols, rows);
    else
        tile_iterate<false>(r, ex, iters, k, lx, tr, gxb, gyb, cols, rows);

    // ---- write the centre tile back through the staging buffer ----
    {
        const int o0 = (2 * tr) * R + 4 * lx, o1 = o0 + R;
        st4(stage + 0 * PLANE_F + o0, r.u1[0]);  st4(stage + 0 * PLANE_F + o1, r.u1[1]);
        st4(stage + 1 * PLANE_F + o0, r.u2[0]);  st4(stage + 1 * PLANE_F + o1, r.u2[1]);
        st4(stage + 2 * PLANE_F + o0, r.p11[0]); st4(stage + 2 * PLANE_F + o1, r.p11[1]);
        st4(stage + 3 * PLANE_F + o0, r.p12[0]); st4(stage + 3 * PLANE_F + o1, r.p12[1]);
        st4(stage + 4 * PLANE_F + o0, r.p21[0]); st4(stage + 4 * PLANE_F + o1, r.p21[1]);
        st4(stage + 5 * PLANE_F + o0, r.p22[0]); st4(stage + 5 * PLANE_F + o1, r.p22[1]);
    }
    __syncthreads();
    const int tw = min(tile, cols - (int)blockIdx.x * tile);
    const int th = min(tile, rows - (int)blockIdx.y * tile);
    const int n_out = tw * th;
#pragma unroll 1
    for (int pl = 0; pl < N_OUT; ++pl) {
        const Plane P = out.p[pl];
        const float *src = stage + pl * PLANE_F;
        for (int idx = tid; idx < n_out; idx += NT) {
            const int ty = idx / tw, tx = idx - ty * tw;
            P.at(gy0 + iters + ty, gx0 + iters + tx) = src[(iters + ty) * R + iters + tx];
        }
    }
}

}  // namespace

cudaError_t tvl1_blocked_init() {
    static bool done[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64 && done[dev]) return cudaSuccess;
    e = cudaFuncSetAttribute(k_tvl1_blocked, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    if (e == cudaSuccess && dev >= 0 && dev < 64) done[dev] = true;
    return e;
}

int tvl1_blocked_pick_k(int knob, int remaining, int rows, int cols) {
    (void)rows;
    (void)cols;
    int kk = knob > 0 ? knob : 5;
    if (kk > TVL1_KMAX) kk = TVL1_KMAX;
    if (kk > remaining) kk = remaining;
    return kk;
}

void tvl1_blocked_launch(Ctx &c, int cls, const Tvl1BlockedPlanes &B, int cur, int rows, int cols,
                         const Tvl1Scalars &k, int iters) {
    const Tvl1State &si = B.s[cur], &so = B.s[cur ^ 1];
    InPlanes in;
    in.p[0] = B.I1wx; in.p[1] = B.I1wy; in.p[2] = B.grad; in.p[3] = B.rho_c;
    in.p[4] = si.u1; in.p[5] = si.u2; in.p[6] = si.p11; in.p[7] = si.p12; in.p[8] = si.p21; in.p[9] = si.p22;
    OutPlanes out;
    out.p[0] = so.u1; out.p[1] = so.u2; out.p[2] = so.p11; out.p[3] = so.p12; out.p[4] = so.p21; out.p[5] = so.p22;
    const int tile = R - 2 * iters;
    const dim3 grid(div_up(cols, tile), div_up(rows, tile));
    const double bytes = 64.0 * (double)rows * cols * iters;  // algorithmic: 64 B / px / iteration
    B2F_LAUNCH(c, cls, bytes, k_tvl1_blocked, grid, dim3(NT), SMEM_BYTES, in, out, rows, cols, k, iters, tile);
}

}  // namespace b2f
