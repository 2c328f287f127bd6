// tvl1_math.cuh -- the per-pixel arithmetic of the Dual TV-L1 inner loop, written once with
// explicit rounding intrinsics so that every kernel variant (unfused, fused, temporally blocked)
// produces bit-identical values for the same inputs regardless of how ptxas would otherwise
// contract mul+add pairs.  Semantics follow the reference kernels:
//   estimateUKernel             modules/cudaoptflow/src/cuda/tvl1flow.cu:209-288
//   estimateDualVariablesKernel modules/cudaoptflow/src/cuda/tvl1flow.cu:313-348
// Deliberate arithmetic choices (documented in DESIGN.md, covered by the tolerance tests):
//   * -rho/grad is -rho * rcp.approx(grad) (MUFU.RCP, <= 2 ulp) instead of an IEEE divide, and the
//     three-way threshold test is written as the equivalent clamp (see tvl1_threshold);
//   * hypotf(a,b) is sqrt.approx(a*a + b*b) and the dual normalisation multiplies by a shared
//     rcp.approx instead of IEEE divides (see tvl1_update_p2);
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

// Per-pixel constant of the thresholding step: 1/|grad I|^2, or a huge value where the reference
// treats the gradient as zero (grad <= FLT_EPSILON), so that the clamp below degenerates to the
// reference's sign test there.  Constant over the inner iterations of a warp: the warp kernels
// evaluate it once per pixel and store it in the `grad` plane.
__device__ __forceinline__ float tvl1_inv_grad(float grad) { return grad > FLT_EPSILON ? rcp_approx(grad) : 1e30f; }

// Thresholding step TH (tvl1flow.cu:236-262): the multiplier fi with d = fi * (Ix, Iy, gamma).
// The reference's three-way test  rho < -l_t*g -> l_t ; rho > l_t*g -> -l_t ; else -rho/g  is the
// clamp of -rho/g to [-l_t, l_t] (TH is continuous in rho), evaluated branch-free with two FMNMX.
__device__ __forceinline__ float tvl1_threshold(float rho, float inv_grad, float l_t) {
    return fminf(fmaxf(__fmul_rn(-rho, inv_grad), -l_t), l_t);
}

// One primal update for a single pixel.  pl = p11(x-1) (0 at x==0), pu = p12(y-1) (0 at y==0).
__device__ __forceinline__ void tvl1_update_u(const Tvl1Scalars &k, float Ix, float Iy, float inv_grad, float rho_c,
                                              float u1, float u2, float p11, float p11_l, float p12, float p12_u,
                                              float p21, float p21_l, float p22, float p22_u, float &u1n,
                                              float &u2n) {
    const float rho = __fmaf_rn(Iy, u2, __fmaf_rn(Ix, u1, rho_c));  // two FMAs (the reference's nvcc build contracts too)
    const float fi = tvl1_threshold(rho, inv_grad, k.l_t);
    const float v1 = __fmaf_rn(fi, Ix, u1);
    const float v2 = __fmaf_rn(fi, Iy, u2);
    const float div1 = __fadd_rn(__fsub_rn(p11, p11_l), __fsub_rn(p12, p12_u));
    const float div2 = __fadd_rn(__fsub_rn(p21, p21_l), __fsub_rn(p22, p22_u));
    u1n = __fmaf_rn(k.theta, div1, v1);
    u2n = __fmaf_rn(k.theta, div2, v2);
}

// Dual update of both flow components of one pixel (estimateDualVariablesKernel, tvl1flow.cu:313-348).
// ux*, uy* are the forward differences (0 on the last column / row).  The two normalisations
// 1/(1 + taut*g1), 1/(1 + taut*g2) share ONE reciprocal: r = 1/(a1*a2), inv1 = r*a2, inv2 = r*a1 (a1, a2 >= 1).
// The SFU path (16 lanes/SM on B200, fed through the MIO queue together with LDS / STS / SHFL) is the first thing
// to saturate in the dual half iteration: measured in round 2, a fourth MUFU per pixel instead of these three FMULs
// makes the kernel 6 % SLOWER (mio_throttle stalls 0.86 -> 1.49 per issue, ncu) although it executes 6 % fewer
// instructions.
__device__ __forceinline__ void tvl1_update_p2(float taut, float ux1, float uy1, float ux2, float uy2, float &p11,
                                               float &p12, float &p21, float &p22) {
    const float g1 = sqrt_approx(__fmaf_rn(ux1, ux1, __fmul_rn(uy1, uy1)));
    const float g2 = sqrt_approx(__fmaf_rn(ux2, ux2, __fmul_rn(uy2, uy2)));
    const float a1 = __fmaf_rn(taut, g1, 1.0f);
    const float a2 = __fmaf_rn(taut, g2, 1.0f);
    const float r = rcp_approx(__fmul_rn(a1, a2));
    const float inv1 = __fmul_rn(r, a2), inv2 = __fmul_rn(r, a1);
    p11 = __fmul_rn(__fmaf_rn(taut, ux1, p11), inv1);
    p12 = __fmul_rn(__fmaf_rn(taut, uy1, p12), inv1);
    p21 = __fmul_rn(__fmaf_rn(taut, ux2, p21), inv2);
    p22 = __fmul_rn(__fmaf_rn(taut, uy2, p22), inv2);
}

// Single-component dual update (third, illumination component when gamma != 0).
__device__ __forceinline__ void tvl1_update_p(float taut, float ux, float uy, float &pa, float &pb) {
    const float g = sqrt_approx(__fmaf_rn(ux, ux, __fmul_rn(uy, uy)));
    const float inv = rcp_approx(__fmaf_rn(taut, g, 1.0f));
    pa = __fmul_rn(__fmaf_rn(taut, ux, pa), inv);
    pb = __fmul_rn(__fmaf_rn(taut, uy, pb), inv);
}

// ---------------------------------------------------------------------------------------------
// Packed FP32 (Blackwell fma.rn.f32x2 / add.rn.f32x2 / mul.rn.f32x2 -> FFMA2 / FADD2 / FMUL2): the same
// arithmetic on TWO independent pixels per instruction.  Every lane is an IEEE round-to-nearest
// operation in the same order as the scalar functions above, so results are bit-identical to them;
// the FP32 work of the iteration kernel takes half the issue slots.
// ---------------------------------------------------------------------------------------------
typedef float2 f2;
__device__ __forceinline__ unsigned long long f2_pack(f2 a) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y));
    return r;
}
__device__ __forceinline__ f2 f2_unpack(unsigned long long a) {
    f2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(a));
    return r;
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(f2_pack(a)), "l"(f2_pack(b)), "l"(f2_pack(c)));
    return f2_unpack(d);
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_pack(a)), "l"(f2_pack(b)));
    return f2_unpack(d);
}
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_pack(a)), "l"(f2_pack(b)));
    return f2_unpack(d);
}
__device__ __forceinline__ f2 sub2(f2 a, f2 b) {
    unsigned long long d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_pack(a)), "l"(f2_pack(b)));
    return f2_unpack(d);
}
__device__ __forceinline__ f2 splat2(float v) { return make_float2(v, v); }

// tvl1_update_u for two pixels.  ninv = -tvl1_inv_grad(grad): (-rho) * inv == rho * (-inv) bit for bit.
__device__ __forceinline__ void tvl1_update_u_x2(const Tvl1Scalars &k, f2 Ix, f2 Iy, f2 ninv, f2 rho_c, f2 u1, f2 u2,
                                                 f2 p11, f2 p11_l, f2 p12, f2 p12_u, f2 p21, f2 p21_l, f2 p22,
                                                 f2 p22_u, f2 &u1n, f2 &u2n) {
    const f2 rho = fma2(Iy, u2, fma2(Ix, u1, rho_c));
    f2 fi = mul2(rho, ninv);
    fi.x = fminf(fmaxf(fi.x, -k.l_t), k.l_t);
    fi.y = fminf(fmaxf(fi.y, -k.l_t), k.l_t);
    const f2 v1 = fma2(fi, Ix, u1);
    const f2 v2 = fma2(fi, Iy, u2);
    const f2 div1 = add2(sub2(p11, p11_l), sub2(p12, p12_u));
    const f2 div2 = add2(sub2(p21, p21_l), sub2(p22, p22_u));
    const f2 th = splat2(k.theta);
    u1n = fma2(th, div1, v1);
    u2n = fma2(th, div2, v2);
}

// tvl1_update_p2 for two pixels (the three MUFU per pixel stay scalar: there is no packed SFU op).
__device__ __forceinline__ void tvl1_update_p2_x2(float taut, f2 ux1, f2 uy1, f2 ux2, f2 uy2, f2 &p11, f2 &p12,
                                                  f2 &p21, f2 &p22) {
    const f2 s1 = fma2(ux1, ux1, mul2(uy1, uy1));
    const f2 s2 = fma2(ux2, ux2, mul2(uy2, uy2));
    const f2 g1 = make_float2(sqrt_approx(s1.x), sqrt_approx(s1.y));
    const f2 g2 = make_float2(sqrt_approx(s2.x), sqrt_approx(s2.y));
    const f2 t = splat2(taut), one = splat2(1.0f);
    const f2 a1 = fma2(t, g1, one);
    const f2 a2 = fma2(t, g2, one);
    const f2 m = mul2(a1, a2);
    const f2 r = make_float2(rcp_approx(m.x), rcp_approx(m.y));
    const f2 inv1 = mul2(r, a2), inv2 = mul2(r, a1);
    p11 = mul2(fma2(t, ux1, p11), inv1);
    p12 = mul2(fma2(t, uy1, p12), inv1);
    p21 = mul2(fma2(t, ux2, p21), inv2);
    p22 = mul2(fma2(t, uy2, p22), inv2);
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
