// tvl1_math.cuh -- the per-pixel arithmetic of the Dual TV-L1 inner loop, written once with
// explicit rounding intrinsics so that every kernel variant (unfused, fused, temporally blocked)
// produces bit-identical values for the same inputs regardless of how ptxas would otherwise
// contract mul+add pairs.  Semantics follow the reference kernels:
//   estimateUKernel             modules/cudaoptflow/src/cuda/tvl1flow.cu:209-288
//   estimateDualVariablesKernel modules/cudaoptflow/src/cuda/tvl1flow.cu:313-348
// Deliberate arithmetic choices (documented in DESIGN.md, covered by the tolerance tests):
//   * -rho/grad is -rho * rcp.approx(grad) (MUFU.RCP, <= 2 ulp) instead of an IEEE divide;
//   * hypotf(a,b) is sqrt.approx(a*a + b*b) and the dual normalisation multiplies by
//     rcp.approx(1 + taut*g) instead of two IEEE divides;
//   * the divergence at the first row/column uses a zero ghost value, (p - 0) + (q - q_up), where
//     the reference writes p + q - q_up: same value up to one rounding on that column only.
#pragma once
#include "common.cuh"

namespace b2f {

struct Tvl1Scalars {
    float l_t;    // lambda * theta
    float theta;  // theta
    float taut;   // tau / theta
    float gamma;  // illumination weight (0 = off)
};

#ifdef __CUDACC__

// Thresholding step TH: returns the multiplier fi such that d = fi * (Ix, Iy, gamma).
// Branch-free (selects only): a per-pixel divergent branch around the reciprocal serialises the
// eight pixels a thread owns in the blocked kernel.
__device__ __forceinline__ float tvl1_threshold(float rho, float grad, float l_t) {
    const float lg = __fmul_rn(l_t, grad);
    const float q = __fmul_rn(-rho, rcp_approx(grad));  // garbage when grad == 0, discarded below
    float fi = grad > FLT_EPSILON ? q : 0.f;
    fi = rho > lg ? -l_t : fi;
    fi = rho < -lg ? l_t : fi;
    return fi;
}

// One primal update for a single pixel.  pl = p11(x-1) (0 at x==0), pu = p12(y-1) (0 at y==0).
__device__ __forceinline__ void tvl1_update_u(const Tvl1Scalars &k, float Ix, float Iy, float grad, float rho_c,
                                              float u1, float u2, float p11, float p11_l, float p12, float p12_u,
                                              float p21, float p21_l, float p22, float p22_u, float &u1n,
                                              float &u2n) {
    const float rho = __fadd_rn(rho_c, __fmaf_rn(Iy, u2, __fmul_rn(Ix, u1)));
    const float fi = tvl1_threshold(rho, grad, k.l_t);
    const float v1 = __fmaf_rn(fi, Ix, u1);
    const float v2 = __fmaf_rn(fi, Iy, u2);
    const float div1 = __fadd_rn(__fsub_rn(p11, p11_l), __fsub_rn(p12, p12_u));
    const float div2 = __fadd_rn(__fsub_rn(p21, p21_l), __fsub_rn(p22, p22_u));
    u1n = __fmaf_rn(k.theta, div1, v1);
    u2n = __fmaf_rn(k.theta, div2, v2);
}

// One dual update for one flow component.  ux, uy are the forward differences (0 on the last
// column / row).
__device__ __forceinline__ void tvl1_update_p(float taut, float ux, float uy, float &pa, float &pb) {
    const float g = sqrt_approx(__fmaf_rn(ux, ux, __fmul_rn(uy, uy)));
    const float inv = rcp_approx(__fmaf_rn(taut, g, 1.0f));
    pa = __fmul_rn(__fmaf_rn(taut, ux, pa), inv);
    pb = __fmul_rn(__fmaf_rn(taut, uy, pb), inv);
}

// Keys bicubic kernel, a = -0.5 (tvl1flow.cu:89-104).
__device__ __forceinline__ float bicubic_coeff(float x_) {
    const float x = fabsf(x_);
    if (x <= 1.0f) return x * x * (1.5f * x - 2.5f) + 1.0f;
    if (x < 2.0f) return x * (x * (-0.5f * x + 2.5f) - 4.0f) + 2.0f;
    return 0.0f;
}

#endif  // __CUDACC__

}  // namespace b2f
