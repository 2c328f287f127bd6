// brox.cu -- cv::cuda::BroxOpticalFlow re-implemented for sm_100a.
//
// Reference being replaced:
//   modules/cudaoptflow/src/brox.cpp:129-188                       (BroxOpticalFlowImpl::calc shim)
//   modules/cudalegacy/src/cuda/NCVBroxOpticalFlow.cu:188-554,598-985 (solver)         [NB]
//   modules/cudalegacy/src/cuda/NPP_staging.cu:1433-1501,2072-2231     (filters, resizers) [NS]
//
// Differences by design:
//  * one stream, no host synchronisation, whole solve captured in a CUDA graph (the reference issues
//    >= 6 cudaStreamSynchronize per call plus one per pyramid level and runs its filters/resizers on
//    the global NPP stream regardless of the caller's, NS:60-75);
//  * no texture objects: bilinear sampling with mirror addressing is done in software with exact
//    float32 weights (the reference's hardware filtering uses 9-bit fixed-point weights);
//  * prepare_sor stage 1 + stage 2 fused (the neighbours' diffusivities are recomputed instead of
//    making a second pass), u += du fused into the prolongation.
//
// Kernel classes: 0 sor (red/black half sweep, 52 B/px algorithmic per pair of half sweeps),
// 1 prepare (104 B/px), 2 derivatives, 3 pyramid (supersample / bicubic prolongation / add).
#include "common.cuh"

#include <cmath>
#include <cstring>
#include <new>
#include <vector>

namespace b2f {

namespace {

enum { CLS_SOR = 0, CLS_PREP = 1, CLS_DERIV = 2, CLS_PYR = 3 };
constexpr float EPS2 = 1e-6f;  // NB:78

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int mirror_filter(int i, int n) {  // NS:1433-1447 (asymmetric low side)
    if (i < 0) i = 1 - i;
    if (i >= n) i = n + n - i - 1;
    return clampi(i, 0, n - 1);
}
__device__ __forceinline__ int mirror_tex(int i, int n) {  // cudaAddressModeMirror on texel indices
    const int p = 2 * n;
    i %= p;
    if (i < 0) i += p;
    return i >= n ? p - 1 - i : i;
}
__device__ __forceinline__ int mirror_load(int i, int n) {  // load_array_element, NB:243-246
    i = max(i, -i - 1);
    i = min(i, n - i + n - 1);
    return clampi(i, 0, n - 1);
}

// Normalised-coordinate, linear, mirror-addressed fetch (software replacement of the texture unit).
__device__ __forceinline__ float tex_bilinear(const Plane &P, int h, int w, float xn, float yn) {
    const float xb = xn * (float)w - 0.5f, yb = yn * (float)h - 0.5f;
    const float fx0 = floorf(xb), fy0 = floorf(yb);
    const float ax = xb - fx0, ay = yb - fy0;
    // keep the integer conversion safe for wild flows; mirror addressing is periodic anyway
    const int x0 = (int)fminf(fmaxf(fx0, -1.0e6f), 1.0e6f), y0 = (int)fminf(fmaxf(fy0, -1.0e6f), 1.0e6f);
    const int xa = mirror_tex(x0, w), xc = mirror_tex(x0 + 1, w);
    const int ya = mirror_tex(y0, h), yc = mirror_tex(y0 + 1, h);
    const float top = __ldg(&P.at(ya, xa)) * (1.f - ax) + __ldg(&P.at(ya, xc)) * ax;
    const float bot = __ldg(&P.at(yc, xa)) * (1.f - ax) + __ldg(&P.at(yc, xc)) * ax;
    return top * (1.f - ay) + bot * ay;
}

// ---------------------------------------------------------------------------------------------
// pyramid: supersample restriction (NS:2073-2149), both frames per launch
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float ss_line(const float *row, int sw, float xmin, float xmax, int ixmin, int ixmax,
                                         float fxmin, float cxmax) {
    float wsum = 1.0f - xmin + fxmin;
    float sum = __ldg(row + ixmin) * (1.0f - xmin + fxmin);
    int sp = ixmin + 1;
    for (int ix = ixmin + 1; ix < ixmax; ++ix) {
        sum += __ldg(row + min(sp, sw - 1));
        sp++;
        wsum += 1.0f;
    }
    sum += __ldg(row + min(sp, sw - 1)) * (cxmax - xmax);
    wsum += cxmax - xmax;
    return sum / wsum;
}

__global__ void __launch_bounds__(256) k_brox_supersample(Plane s0, Plane s1, int sh, int sw, Plane d0, Plane d1, int dh,
                                                          int dw, float scale) {
    const int ix = blockIdx.x * blockDim.x + threadIdx.x;
    const int iy = blockIdx.y * blockDim.y + threadIdx.y;
    if (ix >= dw || iy >= dh) return;
    const Plane src = blockIdx.z ? s1 : s0;
    const float rw = (float)sw, rh = (float)sh;
    const float x = scale * (float)ix, y = scale * (float)iy;
    const float xBegin = fmaxf(x - scale, 0.0f), xEnd = fminf(x + scale, rw - 1.0f);
    const float yBegin = fmaxf(y - scale, 0.0f), yEnd = fminf(y + scale, rh - 1.0f);
    const float fxb = floorf(xBegin), cxe = ceilf(xEnd);
    const int iXBegin = (int)fxb, iXEnd = (int)cxe;
    const float fyb = floorf(yBegin), cye = ceilf(yEnd);
    const int iYBegin = (int)fyb, iYEnd = (int)cye;
    int ry = iYBegin;
    float wsum = 1.0f - yBegin + fyb;
    float sum = ss_line(src.row(ry), sw, xBegin, xEnd, iXBegin, iXEnd, fxb, cxe) * (1.0f - yBegin + fyb);
    ry++;
    for (int yy = iYBegin + 1; yy < iYEnd; ++yy) {
        sum += ss_line(src.row(min(ry, sh - 1)), sw, xBegin, xEnd, iXBegin, iXEnd, fxb, cxe);
        ry++;
        wsum += 1.0f;
    }
    sum += ss_line(src.row(min(ry, sh - 1)), sw, xBegin, xEnd, iXBegin, iXEnd, fxb, cxe) * (cye - yEnd);
    wsum += cye - yEnd;
    (blockIdx.z ? d1 : d0).at(iy, ix) = sum / wsum;
}

// ---------------------------------------------------------------------------------------------
// derivative bank: {1,-8,0,8,-1}/12 with the reference's mirror rule (NS:1449-1501; NB:843-868)
// pass 1: Ix0, Iy0 (from I0), Ix, Iy (from I1);  pass 2: Ixx = dx(Ix), Iyy = dy(Iy), Ixy = dx(Iy)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float d5_row(const Plane &P, int y, int x, int w) {
    float sum = 0.0f;
    sum += __ldg(&P.at(y, mirror_filter(x - 2, w))) * 1.0f;
    sum += __ldg(&P.at(y, mirror_filter(x - 1, w))) * -8.0f;
    sum += __ldg(&P.at(y, mirror_filter(x, w))) * 0.0f;
    sum += __ldg(&P.at(y, mirror_filter(x + 1, w))) * 8.0f;
    sum += __ldg(&P.at(y, mirror_filter(x + 2, w))) * -1.0f;
    return sum * (1.0f / 12.0f);
}
__device__ __forceinline__ float d5_col(const Plane &P, int y, int x, int h) {
    float sum = 0.0f;
    sum += __ldg(&P.at(mirror_filter(y - 2, h), x)) * 1.0f;
    sum += __ldg(&P.at(mirror_filter(y - 1, h), x)) * -8.0f;
    sum += __ldg(&P.at(mirror_filter(y, h), x)) * 0.0f;
    sum += __ldg(&P.at(mirror_filter(y + 1, h), x)) * 8.0f;
    sum += __ldg(&P.at(mirror_filter(y + 2, h), x)) * -1.0f;
    return sum * (1.0f / 12.0f);
}

__global__ void __launch_bounds__(256) k_brox_deriv1(Plane I0, Plane I1, Plane Ix0, Plane Iy0, Plane Ix, Plane Iy,
                                                     int h, int w) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    Ix0.at(y, x) = d5_row(I0, y, x, w);
    Iy0.at(y, x) = d5_col(I0, y, x, h);
    Ix.at(y, x) = d5_row(I1, y, x, w);
    Iy.at(y, x) = d5_col(I1, y, x, h);
}

__global__ void __launch_bounds__(256) k_brox_deriv2(Plane Ix, Plane Iy, Plane Ixx, Plane Iyy, Plane Ixy, int h, int w) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    Ixx.at(y, x) = d5_row(Ix, y, x, w);
    Iyy.at(y, x) = d5_col(Iy, y, x, h);
    Ixy.at(y, x) = d5_row(Iy, y, x, w);
}

// ---------------------------------------------------------------------------------------------
// prepare_sor: stage 1 (NB:340-406) + stage 2 (NB:416-473) fused
// ---------------------------------------------------------------------------------------------
struct BroxLevelPlanes {
    Plane I0, I1, Ix, Ixx, Ix0, Iy, Iyy, Iy0, Ixy;
    Plane u, v, du, dv;
    Plane sx, sy, inv_u, inv_v, num_dudv, num_u, num_v;
};

// diffusivity between (i,j) and (i-1,j)  (NB:188-204)
__device__ __forceinline__ float brox_sx(const Plane &u, const Plane &v, const Plane &du, const Plane &dv, int h, int w,
                                         int j, int i) {
    auto L = [&](const Plane &P, int y, int x) { return __ldg(&P.at(mirror_load(y, h), mirror_load(x, w))); };
    const float u_x = L(u, j, i) + L(du, j, i) - L(u, j, i - 1) - L(du, j, i - 1);
    const float v_x = L(v, j, i) + L(dv, j, i) - L(v, j, i - 1) - L(dv, j, i - 1);
    const float u_y = 0.25f * (L(u, j + 1, i) + L(du, j + 1, i) + L(u, j + 1, i - 1) + L(du, j + 1, i - 1) -
                               L(u, j - 1, i) - L(du, j - 1, i) - L(u, j - 1, i - 1) - L(du, j - 1, i - 1));
    const float v_y = 0.25f * (L(v, j + 1, i) + L(dv, j + 1, i) + L(v, j + 1, i - 1) + L(dv, j + 1, i - 1) -
                               L(v, j - 1, i) - L(dv, j - 1, i) - L(v, j - 1, i - 1) - L(dv, j - 1, i - 1));
    return 0.5f / sqrtf(u_x * u_x + v_x * v_x + u_y * u_y + v_y * v_y + EPS2);
}
// diffusivity between (i,j) and (i,j-1)  (NB:216-227)
__device__ __forceinline__ float brox_sy(const Plane &u, const Plane &v, const Plane &du, const Plane &dv, int h, int w,
                                         int j, int i) {
    auto L = [&](const Plane &P, int y, int x) { return __ldg(&P.at(mirror_load(y, h), mirror_load(x, w))); };
    const float u_y = L(u, j, i) + L(du, j, i) - L(u, j - 1, i) - L(du, j - 1, i);
    const float v_y = L(v, j, i) + L(dv, j, i) - L(v, j - 1, i) - L(dv, j - 1, i);
    const float u_x = 0.25f * (L(u, j, i + 1) + L(u, j - 1, i + 1) + L(du, j, i + 1) + L(du, j - 1, i + 1) -
                               L(u, j, i - 1) - L(u, j - 1, i - 1) - L(du, j, i - 1) - L(du, j - 1, i - 1));
    const float v_x = 0.25f * (L(v, j, i + 1) + L(v, j - 1, i + 1) + L(dv, j, i + 1) + L(dv, j - 1, i + 1) -
                               L(v, j, i - 1) - L(v, j - 1, i - 1) - L(dv, j, i - 1) - L(dv, j - 1, i - 1));
    return 0.5f / sqrtf(u_x * u_x + v_x * v_x + u_y * u_y + v_y * v_y + EPS2);
}

__global__ void __launch_bounds__(256) k_brox_prepare(BroxLevelPlanes P, int h, int w, float alpha, float gamma) {
    const int ig = blockIdx.x * blockDim.x + threadIdx.x;
    const int jg = blockIdx.y * blockDim.y + threadIdx.y;
    if (ig >= w || jg >= h) return;
    float x = (float)ig + 0.5f, y = (float)jg + 0.5f;
    const float uu = P.u.at(jg, ig), vv = P.v.at(jg, ig);
    const float du = P.du.at(jg, ig), dv = P.dv.at(jg, ig);
    const float wx = (x + uu) / (float)w, wy = (y + vv) / (float)h;
    x /= (float)w;
    y /= (float)h;
    const float Iz = tex_bilinear(P.I1, h, w, wx, wy) - tex_bilinear(P.I0, h, w, x, y);
    const float Ix = tex_bilinear(P.Ix, h, w, wx, wy);
    const float Ixz = Ix - tex_bilinear(P.Ix0, h, w, x, y);
    const float Ixy = tex_bilinear(P.Ixy, h, w, wx, wy);
    const float Ixx = tex_bilinear(P.Ixx, h, w, wx, wy);
    const float Iy = tex_bilinear(P.Iy, h, w, wx, wy);
    const float Iyz = Iy - tex_bilinear(P.Iy0, h, w, x, y);
    const float Iyy = tex_bilinear(P.Iyy, h, w, wx, wy);
    const float q0 = Iz + Ix * du + Iy * dv;
    const float q1 = Ixz + Ixx * du + Ixy * dv;
    const float q2 = Iyz + Ixy * du + Iyy * dv;
    float data_term = 0.5f * rsqrtf(q0 * q0 + gamma * (q1 * q1 + q2 * q2) + EPS2);
    data_term /= alpha;

    // smoothness diffusivities of this pixel and of its right / upper neighbours (stage 2 needs them)
    const float sx = ig == 0 ? 0.f : brox_sx(P.u, P.v, P.du, P.dv, h, w, jg, ig);
    const float sy = jg == 0 ? 0.f : brox_sy(P.u, P.v, P.du, P.dv, h, w, jg, ig);
    const float sxr = ig + 1 < w ? brox_sx(P.u, P.v, P.du, P.dv, h, w, jg, ig + 1) : 0.f;
    const float syu = jg + 1 < h ? brox_sy(P.u, P.v, P.du, P.dv, h, w, jg + 1, ig) : 0.f;

    P.num_dudv.at(jg, ig) = data_term * (Ix * Iy + gamma * Ixy * (Ixx + Iyy));
    P.num_u.at(jg, ig) = data_term * (Ix * Iz + gamma * (Ixx * Ixz + Ixy * Iyz));
    P.num_v.at(jg, ig) = data_term * (Iy * Iz + gamma * (Iyy * Iyz + Ixy * Ixz));
    const float den_u = data_term * (Ix * Ix + gamma * (Ixy * Ixy + Ixx * Ixx));
    const float den_v = data_term * (Iy * Iy + gamma * (Ixy * Ixy + Iyy * Iyy));
    P.sx.at(jg, ig) = sx;
    P.sy.at(jg, ig) = sy;
    const float dsum = sx + sxr + sy + syu;
    P.inv_u.at(jg, ig) = 1.0f / (den_u + dsum);
    P.inv_v.at(jg, ig) = 1.0f / (den_v + dsum);
}

// ---------------------------------------------------------------------------------------------
// one red or black half sweep (NB:479-554), omega = 1.99
// ---------------------------------------------------------------------------------------------
template <int IS_BLACK>
__global__ void __launch_bounds__(256) k_brox_sor(BroxLevelPlanes P, Plane du_in, Plane dv_in, Plane du_out,
                                                  Plane dv_out, int h, int w, float omega) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= w || j >= h) return;
    float du = du_in.at(j, i), dv = dv_in.at(j, i);
    if (((i + j) & 1) == IS_BLACK) {
        const int ir = i < w - 1 ? i + 1 : i, il = i > 0 ? i - 1 : i;
        const int ju = j < h - 1 ? j + 1 : j, jd = j > 0 ? j - 1 : j;
        const float s_left = P.sx.at(j, i), s_down = P.sy.at(j, i);
        const float s_right = i < w - 1 ? P.sx.at(j, ir) : 0.0f;
        const float s_up = j < h - 1 ? P.sy.at(ju, i) : 0.0f;
        const float u = P.u.at(j, i), v = P.v.at(j, i);
        const float ssum = s_left + s_right + s_up + s_down;
        const float numerator_dudv = P.num_dudv.at(j, i);
        const float numerator_u =
            (s_left * (P.u.at(j, il) + du_in.at(j, il)) + s_up * (P.u.at(ju, i) + du_in.at(ju, i)) +
             s_right * (P.u.at(j, ir) + du_in.at(j, ir)) + s_down * (P.u.at(jd, i) + du_in.at(jd, i)) - u * ssum -
             P.num_u.at(j, i) - numerator_dudv * dv);
        du = (1.0f - omega) * du + omega * P.inv_u.at(j, i) * numerator_u;
        const float numerator_v =
            (s_left * (P.v.at(j, il) + dv_in.at(j, il)) + s_up * (P.v.at(ju, i) + dv_in.at(ju, i)) +
             s_right * (P.v.at(j, ir) + dv_in.at(j, ir)) + s_down * (P.v.at(jd, i) + dv_in.at(jd, i)) - v * ssum -
             P.num_v.at(j, i) - numerator_dudv * du);
        dv = (1.0f - omega) * dv + omega * P.inv_v.at(j, i) * numerator_v;
    }
    du_out.at(j, i) = du;
    dv_out.at(j, i) = dv;
}

// ---------------------------------------------------------------------------------------------
// Fused red-black SOR: `iters` full iterations (2*iters half sweeps) per launch on a 64x64 region held in
// shared memory.  Each half sweep consumes one halo cell per side, so the centre (64 - 4*iters)^2 tile
// is exact (bit-identical to `iters` pairs of k_brox_sor launches: a colour's update reads only the
// other colour, so updating in place equals the reference's ping-pong).  A level that fits in one region
// needs no halo at all and runs all solver iterations of an inner step in one launch.
// ---------------------------------------------------------------------------------------------
constexpr int SR = 64;           // region edge
constexpr int S_THREADS = 512;

__global__ void __launch_bounds__(S_THREADS, 1)
    k_brox_sor_fused_pp(BroxLevelPlanes P, Plane du_out, Plane dv_out, int h, int w, float omega, int iters, int halo,
                        int tile) {
    extern __shared__ float ssm[];
    float *s_du = ssm, *s_dv = ssm + SR * SR, *s_u = ssm + 2 * SR * SR, *s_v = ssm + 3 * SR * SR;
    float *s_sx = ssm + 4 * SR * SR, *s_sy = ssm + 5 * SR * SR, *s_iu = ssm + 6 * SR * SR, *s_iv = ssm + 7 * SR * SR;
    float *s_nu = ssm + 8 * SR * SR, *s_nv = ssm + 9 * SR * SR, *s_nd = ssm + 10 * SR * SR;
    const int tid = threadIdx.x;
    const int gx0 = blockIdx.x * tile - halo, gy0 = blockIdx.y * tile - halo;

    for (int idx = tid; idx < SR * SR; idx += S_THREADS) {
        const int ry = idx >> 6, rx = idx & 63;
        const int gy = gy0 + ry, gx = gx0 + rx;
        const bool in = gx >= 0 && gy >= 0 && gx < w && gy < h;
        s_du[idx] = in ? P.du.at(gy, gx) : 0.f;
        s_dv[idx] = in ? P.dv.at(gy, gx) : 0.f;
        s_u[idx] = in ? P.u.at(gy, gx) : 0.f;
        s_v[idx] = in ? P.v.at(gy, gx) : 0.f;
        s_sx[idx] = in ? P.sx.at(gy, gx) : 0.f;
        s_sy[idx] = in ? P.sy.at(gy, gx) : 0.f;
        s_iu[idx] = in ? P.inv_u.at(gy, gx) : 0.f;
        s_iv[idx] = in ? P.inv_v.at(gy, gx) : 0.f;
        s_nu[idx] = in ? P.num_u.at(gy, gx) : 0.f;
        s_nv[idx] = in ? P.num_v.at(gy, gx) : 0.f;
        s_nd[idx] = in ? P.num_dudv.at(gy, gx) : 0.f;
    }
    __syncthreads();

    for (int sweep = 0; sweep < 2 * iters; ++sweep) {
        const int colour = sweep & 1;  // sor_pass<0> first, then sor_pass<1> (NB:915-922)
        for (int c = tid; c < SR * SR / 2; c += S_THREADS) {
            const int ry = c >> 5;
            const int rx = ((c & 31) << 1) + ((ry + gy0 + gx0 + colour) & 1);  // (gx + gy) % 2 == colour
            const int gy = gy0 + ry, gx = gx0 + rx;
            if (gx < 0 || gy < 0 || gx >= w || gy >= h) continue;
            // region-edge cells read clamped region neighbours: their values are stale, but they lie in the halo
            const int idx = ry * SR + rx;
            const int il = (gx > 0 && rx > 0) ? idx - 1 : idx, ir = (gx < w - 1 && rx < SR - 1) ? idx + 1 : idx;
            const int id = (gy > 0 && ry > 0) ? idx - SR : idx, iu = (gy < h - 1 && ry < SR - 1) ? idx + SR : idx;
            const float s_left = s_sx[idx], s_down = s_sy[idx];
            const float s_right = gx < w - 1 ? s_sx[ir] : 0.0f;
            const float s_up = gy < h - 1 ? s_sy[iu] : 0.0f;
            const float u = s_u[idx], v = s_v[idx];
            float du = s_du[idx], dv = s_dv[idx];
            const float ssum = s_left + s_right + s_up + s_down;
            const float numerator_dudv = s_nd[idx];
            const float numerator_u = (s_left * (s_u[il] + s_du[il]) + s_up * (s_u[iu] + s_du[iu]) +
                                       s_right * (s_u[ir] + s_du[ir]) + s_down * (s_u[id] + s_du[id]) - u * ssum -
                                       s_nu[idx] - numerator_dudv * dv);
            du = (1.0f - omega) * du + omega * s_iu[idx] * numerator_u;
            const float numerator_v = (s_left * (s_v[il] + s_dv[il]) + s_up * (s_v[iu] + s_dv[iu]) +
                                       s_right * (s_v[ir] + s_dv[ir]) + s_down * (s_v[id] + s_dv[id]) - v * ssum -
                                       s_nv[idx] - numerator_dudv * du);
            dv = (1.0f - omega) * dv + omega * s_iv[idx] * numerator_v;
            s_du[idx] = du;
            s_dv[idx] = dv;
        }
        __syncthreads();
    }

    for (int idx = tid; idx < tile * tile; idx += S_THREADS) {
        const int ty = idx / tile, tx = idx - ty * tile;
        const int gy = gy0 + halo + ty, gx = gx0 + halo + tx;
        if (gx < w && gy < h) {
            du_out.at(gy, gx) = s_du[(halo + ty) * SR + halo + tx];
            dv_out.at(gy, gx) = s_dv[(halo + ty) * SR + halo + tx];
        }
    }
}

// u += du ; v += dv  (NB:929-931), in place
__global__ void __launch_bounds__(256) k_brox_add(Plane u, Plane v, Plane du, Plane dv, int h, int w) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    u.at(y, x) = u.at(y, x) + du.at(y, x);
    v.at(y, x) = v.at(y, x) + dv.at(y, x);
}

// bicubic prolongation (NS:2172-2231) fused with ScaleVector (NB:952-961); u and v per launch
__device__ __forceinline__ float brox_bicubic_coeff(float x_) {
    const float x = fabsf(x_);
    if (x <= 1.0f) return x * x * (1.5f * x - 2.5f) + 1.0f;
    if (x < 2.0f) return x * (x * (-0.5f * x + 2.5f) - 4.0f) + 2.0f;
    return 0.0f;
}

__global__ void __launch_bounds__(256) k_brox_prolong(Plane su, Plane sv, int sh, int sw, Plane du, Plane dv, int dh,
                                                      int dw, float scale, float mul) {
    const int ix = blockIdx.x * blockDim.x + threadIdx.x;
    const int iy = blockIdx.y * blockDim.y + threadIdx.y;
    if (ix >= dw || iy >= dh) return;
    const Plane src = blockIdx.z ? sv : su;
    const float rw = (float)sw, rh = (float)sh;
    const float x = scale * (float)ix, y = scale * (float)iy;
    const float xmin = fmaxf(ceilf(x - 2.0f), 0.0f), xmax = fminf(floorf(x + 2.0f), rw - 1.0f);
    const float ymin = fmaxf(ceilf(y - 2.0f), 0.0f), ymax = fminf(floorf(y + 2.0f), rh - 1.0f);
    float sum = 0.0f, wsum = 0.0f;
    for (float cy = ymin; cy <= ymax; cy += 1.0f) {
        for (float cx = xmin; cx <= xmax; cx += 1.0f) {
            float wx = brox_bicubic_coeff(x - cx);
            const float wy = brox_bicubic_coeff(y - cy);
            wx *= wy;
            sum += wx * __ldg(&src.at((int)cy, (int)cx));
            wsum += wx;
        }
    }
    const float r = (!wsum) ? 0.f : sum / wsum;
    (blockIdx.z ? dv : du).at(iy, ix) = r * mul;
}

// ---------------------------------------------------------------------------------------------
// host engine
// ---------------------------------------------------------------------------------------------
struct BLevel {
    int rows = 0, cols = 0;
    Plane I0, I1;
};

class BroxEngine : public b2f_handle {
public:
    explicit BroxEngine(const b2f_brox_params &p) : P(p) { algo = ALGO_BROX; }
    ~BroxEngine() override { destroy_graph(); }
    b2f_brox_params P;

    int calc(const b2f_image *I0, const b2f_image *I1, b2f_image *flow, cudaStream_t s) override;
    int set_param(int id, double v) override;
    int get_param(int id, double *v) const override;
    const char *default_name() const override { return "DenseOpticalFlow.BroxOpticalFlow"; }
    const char *class_name(int cls) const override {
        static const char *n[] = {"brox_sor", "brox_prepare", "brox_derivatives", "brox_pyramid"};
        return (cls >= 0 && cls < 4) ? n[cls] : "";
    }
    size_t workspace_bytes(int rows, int cols, int type) override {
        (void)type;
        Layout L;
        return layout(rows, cols, true, L);
    }

private:
    struct Layout {
        int rows = 0, cols = 0;
        b2f_brox_params P{};
        std::vector<BLevel> levels;
        float *shared[20] = {};
    };
    Layout L_;
    enum { S_IX = 0, S_IXX, S_IX0, S_IY, S_IYY, S_IY0, S_IXY, S_U, S_V, S_UN, S_VN, S_DU, S_DV, S_DUN, S_DVN, S_SX, S_SY,
           S_INVU, S_INVV, S_NDUDV, S_COUNT };
    float *extra_[2] = {};  // num_u, num_v
    int final_ui_ = 0;      // which (u, v) buffer pair holds the result (fixed by the level count)

    cudaGraphExec_t graph_exec_ = nullptr;
    struct Key { int rows = 0, cols = 0; b2f_brox_params P{}; void *base = nullptr; } key_;
    uint64_t g_launches_ = 0, g_cls_launches_[B2F_MAX_KERNEL_CLASSES] = {};
    double g_cls_bytes_[B2F_MAX_KERNEL_CLASSES] = {};
    int g_iters_ = 0;
    void destroy_graph() {
        if (graph_exec_) cudaGraphExecDestroy(graph_exec_);
        graph_exec_ = nullptr;
    }

    size_t layout(int rows, int cols, bool counting, Layout &L);
    cudaError_t ensure_workspace(int rows, int cols);
    Plane sp(int idx, int cols) const { return Plane{L_.shared[idx], plane_pitch(cols)}; }
    void solve(Ctx &c);
};

size_t BroxEngine::layout(int rows, int cols, bool counting, Layout &L) {
    Arena tmp;
    Arena &A = counting ? tmp : arena;
    A.begin(counting);
    L.rows = rows;
    L.cols = cols;
    L.P = P;
    L.levels.clear();
    // pyramid sizes, NB:730-785: ceilf(src * scale) with a cumulative float scale
    const float sf = static_cast<float>(P.scale_factor);
    float scale = 1.0f * sf;
    BLevel l0;
    l0.rows = rows;
    l0.cols = cols;
    l0.I0 = A.plane(rows, cols);
    l0.I1 = A.plane(rows, cols);
    L.levels.push_back(l0);
    int pw = cols, ph = rows;
    while (pw > 15 && ph > 15 && (int)L.levels.size() < P.outer_iterations) {
        BLevel lv;
        lv.cols = static_cast<int>(ceilf(cols * scale));
        lv.rows = static_cast<int>(ceilf(rows * scale));
        lv.I0 = A.plane(lv.rows, lv.cols);
        lv.I1 = A.plane(lv.rows, lv.cols);
        L.levels.push_back(lv);
        scale *= sf;
        pw = lv.cols;
        ph = lv.rows;
    }
    for (int i = 0; i < S_COUNT; ++i) L.shared[i] = A.plane(rows, cols).p;
    float *a = A.plane(rows, cols).p, *b = A.plane(rows, cols).p;
    if (!counting) {
        extra_[0] = a;
        extra_[1] = b;
    }
    return A.used();
}

cudaError_t BroxEngine::ensure_workspace(int rows, int cols) {
    if (L_.rows == rows && L_.cols == cols && std::memcmp(&L_.P, &P, sizeof(P)) == 0 && arena.capacity() > 0)
        return cudaSuccess;
    Layout tmp;
    const size_t need = layout(rows, cols, true, tmp);
    destroy_graph();
    cudaError_t e = arena.reserve(need);
    if (e != cudaSuccess) return e;
    layout(rows, cols, false, L_);
    return cudaSuccess;
}

void BroxEngine::solve(Ctx &c) {
    const dim3 block(32, 8);
    const size_t sor_smem = sizeof(float) * 11 * SR * SR;
    const float inv_sf = 1.0f / static_cast<float>(P.scale_factor);  // kernels get 1/xFactor (NS:2264,2270)
    const int nl = static_cast<int>(L_.levels.size());
    // restriction, level by level from the previous one
    for (int l = 1; l < nl; ++l) {
        const BLevel &a = L_.levels[l - 1], &b = L_.levels[l];
        const dim3 grid(div_up(b.cols, 32), div_up(b.rows, 8), 2);
        B2F_LAUNCH(c, CLS_PYR, 2.0 * 4.0 * ((double)a.rows * a.cols + (double)b.rows * b.cols), k_brox_supersample, grid,
                   block, 0, a.I0, a.I1, a.rows, a.cols, b.I0, b.I1, b.rows, b.cols, inv_sf);
    }
    int ui = 0;  // (u, v) live in S_U/S_V (0) or S_UN/S_VN (1)
    {
        const BLevel &lc = L_.levels[nl - 1];
        fill_plane(c, sp(S_U, lc.cols), lc.rows, lc.cols, 0.f);
        fill_plane(c, sp(S_V, lc.cols), lc.rows, lc.cols, 0.f);
    }
    for (int l = nl - 1; l >= 0; --l) {
        const BLevel &lv = L_.levels[l];
        const int h = lv.rows, w = lv.cols;
        const double npx = (double)h * w;
        const dim3 grid(div_up(w, 32), div_up(h, 8));
        BroxLevelPlanes T;
        T.I0 = lv.I0; T.I1 = lv.I1;
        T.Ix = sp(S_IX, w); T.Ixx = sp(S_IXX, w); T.Ix0 = sp(S_IX0, w);
        T.Iy = sp(S_IY, w); T.Iyy = sp(S_IYY, w); T.Iy0 = sp(S_IY0, w); T.Ixy = sp(S_IXY, w);
        T.u = sp(ui ? S_UN : S_U, w); T.v = sp(ui ? S_VN : S_V, w);
        T.du = sp(S_DU, w); T.dv = sp(S_DV, w);
        T.sx = sp(S_SX, w); T.sy = sp(S_SY, w); T.inv_u = sp(S_INVU, w); T.inv_v = sp(S_INVV, w);
        T.num_dudv = sp(S_NDUDV, w);
        T.num_u = Plane{extra_[0], plane_pitch(w)};
        T.num_v = Plane{extra_[1], plane_pitch(w)};
        Plane dun = sp(S_DUN, w), dvn = sp(S_DVN, w);

        fill_plane(c, T.du, h, w, 0.f);
        fill_plane(c, T.dv, h, w, 0.f);
        B2F_LAUNCH(c, CLS_DERIV, 24.0 * npx, k_brox_deriv1, grid, block, 0, T.I0, T.I1, T.Ix0, T.Iy0, T.Ix, T.Iy, h, w);
        B2F_LAUNCH(c, CLS_DERIV, 20.0 * npx, k_brox_deriv2, grid, block, 0, T.Ix, T.Iy, T.Ixx, T.Iyy, T.Ixy, h, w);
        for (int in = 0; in < P.inner_iterations; ++in) {
            B2F_LAUNCH(c, CLS_PREP, 104.0 * npx, k_brox_prepare, grid, block, 0, T, h, w, static_cast<float>(P.alpha),
                       static_cast<float>(P.gamma));
            const float omega = 1.99f;  // NB:914
            if (knobs.kernel_path != 1) {
                // fused path: up to 5 full iterations per launch (all of them when the level fits one region).
                // A CTA reads its halo from global memory that neighbouring CTAs overwrite with their centre
                // tiles at the end, so launches ping-pong between (du, dv) and (dun, dvn).
                int left = P.solver_iterations;
                const bool single = w <= SR && h <= SR;  // whole level in one region: no halo, one launch
                Plane cdu = T.du, cdv = T.dv, ndu = dun, ndv = dvn;
                while (left > 0) {
                    const int it = single ? left : (left < 5 ? left : 5);
                    const int halo = single ? 0 : 2 * it;
                    const int tile = single ? SR : SR - 2 * halo;
                    const dim3 gs(single ? 1 : div_up(w, tile), single ? 1 : div_up(h, tile));
                    BroxLevelPlanes Q = T;
                    Q.du = cdu;
                    Q.dv = cdv;
                    B2F_LAUNCH(c, CLS_SOR, 52.0 * npx * it, k_brox_sor_fused_pp, gs, dim3(S_THREADS), sor_smem, Q, ndu, ndv,
                               h, w, omega, it, halo, tile);
                    Plane t1 = cdu; cdu = ndu; ndu = t1;
                    Plane t2 = cdv; cdv = ndv; ndv = t2;
                    left -= it;
                    c.stats->iterations_run += it;
                }
                if (cdu.p != T.du.p) {  // odd number of launches: result sits in the spare pair -> swap roles
                    Plane t1 = T.du; T.du = dun; dun = t1;
                    Plane t2 = T.dv; T.dv = dvn; dvn = t2;
                }
            } else {
                for (int s = 0; s < P.solver_iterations; ++s) {
                    B2F_LAUNCH(c, CLS_SOR, 26.0 * npx, k_brox_sor<0>, grid, block, 0, T, T.du, T.dv, dun, dvn, h, w, omega);
                    B2F_LAUNCH(c, CLS_SOR, 26.0 * npx, k_brox_sor<1>, grid, block, 0, T, dun, dvn, T.du, T.dv, h, w, omega);
                    c.stats->iterations_run++;
                }
            }
        }
        B2F_LAUNCH(c, CLS_PYR, 24.0 * npx, k_brox_add, grid, block, 0, T.u, T.v, T.du, T.dv, h, w);
        if (l > 0) {
            const BLevel &nx = L_.levels[l - 1];
            Plane nu = sp(ui ? S_U : S_UN, nx.cols), nv = sp(ui ? S_V : S_VN, nx.cols);
            const dim3 g2(div_up(nx.cols, 32), div_up(nx.rows, 8), 2);
            // nppiStResize(..., 1/scale_factor, bicubic) -> kernel scale = scale_factor; then ScaleVector(1/scale_factor)
            B2F_LAUNCH(c, CLS_PYR, 2.0 * 4.0 * (npx + (double)nx.rows * nx.cols), k_brox_prolong, g2, block, 0, T.u, T.v, h,
                       w, nu, nv, nx.rows, nx.cols, 1.0f / inv_sf, inv_sf);
            ui ^= 1;
        }
    }
    final_ui_ = ui;
}

int BroxEngine::calc(const b2f_image *I0, const b2f_image *I1, b2f_image *flow, cudaStream_t s) {
    // preconditions: brox.cpp:134-135 (CV_32FC1, equal size/type), NB:606-617
    if (I0->type != B2F_32FC1 || I1->type != B2F_32FC1) return B2F_UNSUPPORTED_TYPE;
    if (I0->rows != I1->rows || I0->cols != I1->cols) return B2F_SIZE_MISMATCH;
    if (!flow_type_ok(flow)) return B2F_UNSUPPORTED_TYPE;
    if (flow->rows != I0->rows || flow->cols != I0->cols) return B2F_SIZE_MISMATCH;
    if (!(P.alpha > 0.0) || P.gamma < 0.0 || P.inner_iterations <= 0 || P.outer_iterations <= 0 ||
        P.solver_iterations <= 0)
        return B2F_BAD_ARG;
    if (!(P.scale_factor > 0.0 && P.scale_factor < 1.0)) return B2F_BAD_ARG;
    if (I0->step < (size_t)I0->cols * 4 || I1->step < (size_t)I1->cols * 4 || flow->step < (size_t)flow->cols * 8)
        return B2F_BAD_ARG;
    const int rows = I0->rows, cols = I0->cols;
    Ctx c = make_ctx(s);
    {
        static bool attr_done[64] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev >= 0 && dev < 64 && !attr_done[dev]) {
            c.check(cudaFuncSetAttribute(k_brox_sor_fused_pp, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(sizeof(float) * 11 * SR * SR)));
            attr_done[dev] = c.ok();
        }
    }
    c.check(ensure_workspace(rows, cols));
    if (!c.ok()) return finish(c, s);
    stats.levels = static_cast<int>(L_.levels.size());
    stats.iterations_run = 0;

    const ImageView v0{I0->data, I0->step, rows, cols, I0->type};
    const ImageView v1{I1->data, I1->step, rows, cols, I1->type};
    const ImageView vf = flow_view(flow, rows, cols);
    convert_pair(c, CLS_PYR, v0, v1, L_.levels[0].I0, L_.levels[0].I1, 1.0f);  // NB:714-718 copy into aligned planes

    const bool want_graph = knobs.use_graph && !profiling && s != nullptr;
    if (want_graph) {
        const bool hit = graph_exec_ && key_.rows == rows && key_.cols == cols && std::memcmp(&key_.P, &P, sizeof(P)) == 0 &&
                         key_.base == L_.levels[0].I0.p;
        if (!hit) {
            destroy_graph();
            cudaStream_t cs = nullptr;
            c.check(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
            if (c.ok()) {
                Ctx g = make_ctx(cs);
                b2f_stats scratch = stats;
                g.stats = &scratch;
                g.capturing = true;
                g.check(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
                if (g.ok()) solve(g);
                cudaGraph_t graph = nullptr;
                g.check(cudaStreamEndCapture(cs, &graph));
                if (g.ok() && graph) g.check(cudaGraphInstantiate(&graph_exec_, graph, 0));
                if (graph) cudaGraphDestroy(graph);
                cudaStreamDestroy(cs);
                g_launches_ = scratch.launches - stats.launches;
                for (int i = 0; i < B2F_MAX_KERNEL_CLASSES; ++i) {
                    g_cls_launches_[i] = scratch.class_launches[i] - stats.class_launches[i];
                    g_cls_bytes_[i] = scratch.class_bytes[i] - stats.class_bytes[i];
                }
                g_iters_ = scratch.iterations_run;
                c.check(g.err);
                if (c.ok()) {
                    key_.rows = rows; key_.cols = cols; key_.P = P; key_.base = L_.levels[0].I0.p;
                } else {
                    destroy_graph();
                }
            }
        }
        if (c.ok() && graph_exec_) {
            c.check(cudaGraphLaunch(graph_exec_, s));
            stats.launches += g_launches_;
            for (int i = 0; i < B2F_MAX_KERNEL_CLASSES; ++i) {
                stats.class_launches[i] += g_cls_launches_[i];
                stats.class_bytes[i] += g_cls_bytes_[i];
            }
            stats.iterations_run = g_iters_;
        }
    } else {
        solve(c);
    }
    merge_flow(c, CLS_PYR, sp(final_ui_ ? S_UN : S_U, cols), sp(final_ui_ ? S_VN : S_V, cols), vf);  // brox.cpp:186-187
    return finish(c, s);
}

int BroxEngine::set_param(int id, double v) {
    switch (id) {
        case B2F_BROX_ALPHA: P.alpha = v; break;
        case B2F_BROX_GAMMA: P.gamma = v; break;
        case B2F_BROX_SCALE_FACTOR: P.scale_factor = v; break;
        case B2F_BROX_INNER_ITERATIONS: P.inner_iterations = static_cast<int>(v); break;
        case B2F_BROX_OUTER_ITERATIONS: P.outer_iterations = static_cast<int>(v); break;
        case B2F_BROX_SOLVER_ITERATIONS: P.solver_iterations = static_cast<int>(v); break;
        default: return B2F_BAD_ARG;
    }
    return B2F_OK;
}

int BroxEngine::get_param(int id, double *v) const {
    switch (id) {
        case B2F_BROX_ALPHA: *v = P.alpha; break;
        case B2F_BROX_GAMMA: *v = P.gamma; break;
        case B2F_BROX_SCALE_FACTOR: *v = P.scale_factor; break;
        case B2F_BROX_INNER_ITERATIONS: *v = P.inner_iterations; break;
        case B2F_BROX_OUTER_ITERATIONS: *v = P.outer_iterations; break;
        case B2F_BROX_SOLVER_ITERATIONS: *v = P.solver_iterations; break;
        default: return B2F_BAD_ARG;
    }
    return B2F_OK;
}

}  // namespace

}  // namespace b2f

extern "C" {

void b2f_brox_default_params(b2f_brox_params *p) {
    if (!p) return;
    p->alpha = 0.197;
    p->gamma = 50.0;
    p->scale_factor = 0.8;
    p->inner_iterations = 5;
    p->outer_iterations = 150;
    p->solver_iterations = 10;
}

int b2f_brox_create(const b2f_brox_params *p, b2f_handle **out) {
    if (!out) return B2F_BAD_ARG;
    b2f_brox_params d;
    b2f_brox_default_params(&d);
    if (p) d = *p;
    *out = new (std::nothrow) b2f::BroxEngine(d);
    return *out ? B2F_OK : B2F_OUT_OF_MEMORY;
}

}  // extern "C"
