// Microbenchmarks behind the round-2 TV-L1 kernel decisions:
//   1. issue / pipe throughput of scalar FFMA vs packed FFMA2 (fma.rn.f32x2) at 4 warps per scheduler,
//      alone and mixed with MUFU, i.e. is the FP32 pipe itself or the issue slot the limit;
//   2. one-way latency of a DSMEM hand-off between two CTAs of a cluster:
//      (a) st.shared::cluster + remote mbarrier.arrive(release.cluster), (b) st.async + complete_tx,
//      (c) barrier.cluster arrive/wait.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_fp32x2_dsmem ubench_fp32x2_dsmem.cu
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdint>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ unsigned long long pk(float2 a) { return *reinterpret_cast<unsigned long long *>(&a); }
__device__ __forceinline__ float2 upk(unsigned long long a) { return *reinterpret_cast<float2 *>(&a); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(pk(a)), "l"(pk(b)), "l"(pk(c)));
    return upk(d);
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk(a)), "l"(pk(b)));
    return upk(d);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk(a)), "l"(pk(b)));
    return upk(d);
}
__device__ __forceinline__ float sqrt_approx(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

constexpr int CH = 8;  // independent chains per thread

// mode 0: scalar FFMA, 2*CH chains (same flop count as mode 1); mode 1: FFMA2, CH chains;
// mode 2: FADD2 + FMUL2 alternating; mode 3: FFMA2 + 1 MUFU per 6 FFMA2 (TV-L1's ratio 3 MUFU : 35 FP32);
// mode 4: scalar FFMA + MUFU same ratio.
template <int MODE>
__global__ void __launch_bounds__(512, 1) k_tput(float *out, int iters, long long *cyc) {
    float2 a[CH], b, c;
    for (int i = 0; i < CH; ++i) a[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f - i);
    b = make_float2(0.999f, 1.001f);
    c = make_float2(1e-3f, -1e-3f);
    float m = 1.5f + threadIdx.x;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 6; ++r) {
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                if (MODE == 0 || MODE == 4) {
                    a[i].x = __fmaf_rn(a[i].x, b.x, c.x);
                    a[i].y = __fmaf_rn(a[i].y, b.y, c.y);
                } else if (MODE == 1 || MODE == 3) {
                    a[i] = fma2(a[i], b, c);
                } else {
                    a[i] = (r & 1) ? add2(a[i], c) : mul2(a[i], b);
                }
            }
            if (MODE == 3 || MODE == 4) {  // per 8 px-pairs' worth ... keep ratio ~ 3 MUFU per 17.5 packed ops
                m = sqrt_approx(m + a[r].x);
                if (r & 1) m = sqrt_approx(m + 2.0f);
            }
        }
    }
    const long long t1 = clock64();
    float s = m;
    for (int i = 0; i < CH; ++i) s += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ---------------- DSMEM hand-off latency ----------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "W1:\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra D1;\n\tbra W1;\n\tD1:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t remote_bar) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void st_cluster(uint32_t addr, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void st_async(uint32_t addr, float v, uint32_t remote_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(addr), "r"(__float_as_uint(v)), "r"(remote_bar) : "memory");
}

// Ping-pong between the two CTAs of a cluster; one warp per CTA takes part (32 lanes x 4 B = one 128-B row).
// mode 0: st.shared::cluster by 32 lanes, each lane arrives (count 32).  mode 1: st.async complete_tx (128 B).
// mode 2: barrier.cluster.arrive.release + wait.acquire by the whole CTA (512 threads).
template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(512, 1) k_pingpong(int rounds, long long *cyc, float *sink) {
    __shared__ __align__(16) float ghost[2][32];
    __shared__ __align__(8) uint64_t bar;
    cg::cluster_group cl = cg::this_cluster();
    const uint32_t rank = cl.block_rank(), peer = rank ^ 1;
    if (threadIdx.x == 0) {
        mbar_init(&bar, MODE == 0 ? 32 : 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cl.sync();
    const uint32_t r_ghost = mapa(smem_u32(&ghost[0][0]), peer);
    const uint32_t r_bar = mapa(smem_u32(&bar), peer);
    float v = threadIdx.x;
    uint32_t parity = 0;
    long long t0 = clock64();
    if (MODE == 2) {
        for (int i = 0; i < rounds; ++i) {
            asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
            asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
        }
    } else if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        if (MODE == 1 && lane == 0) mbar_expect_tx(&bar, 128);
        __syncwarp();
        for (int i = 0; i < rounds; ++i) {
            // even rounds: rank 0 sends, rank 1 receives; odd rounds the other way
            const bool sender = ((i & 1) == (int)rank);
            if (sender) {
                if (MODE == 0) { st_cluster(r_ghost + 4 * lane, v); mbar_arrive_remote(r_bar); }
                else { st_async(r_ghost + 4 * lane, v, r_bar); }
            } else {
                mbar_wait_cluster(&bar, parity);
                parity ^= 1;
                v += ghost[0][lane];
                if (MODE == 1) { if (lane == 0) mbar_expect_tx(&bar, 128); __syncwarp(); }
            }
        }
    }
    long long t1 = clock64();
    cl.sync();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (threadIdx.x < 32) sink[blockIdx.x * 32 + threadIdx.x] = v;
}

template <int MODE>
static int run_tput(const char *name, int sms, double lanes_per_instr) {
    float *out; long long *cyc;
    CK(cudaMalloc(&out, sizeof(float) * sms * 512));
    CK(cudaMalloc(&cyc, sizeof(long long) * sms));
    const int iters = 2000;
    k_tput<MODE><<<sms, 512>>>(out, 10, cyc);
    k_tput<MODE><<<sms, 512>>>(out, iters, cyc);
    CK(cudaDeviceSynchronize());
    long long h[256];
    CK(cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
    double mean = 0; for (int i = 0; i < sms; ++i) mean += h[i]; mean /= sms;
    // FP32 instructions per thread per iteration: 6 * CH (packed) or 6 * CH * 2 (scalar)
    const double instr = (MODE == 0 || MODE == 4) ? 6.0 * CH * 2 : 6.0 * CH;
    const double warp_instr_per_clk_per_smsp = instr * iters * 16 / 4 / mean;
    printf("%-34s cycles %.0f  FP32 warp-instr/clk/SMSP %.3f  fp32 lane-ops/clk/SM %.1f\n", name, mean,
           warp_instr_per_clk_per_smsp, warp_instr_per_clk_per_smsp * 4 * 32 * lanes_per_instr);
    cudaFree(out); cudaFree(cyc);
    return 0;
}

template <int MODE>
static int run_pp(const char *name) {
    long long *cyc; float *sink;
    CK(cudaMalloc(&cyc, sizeof(long long) * 2));
    CK(cudaMalloc(&sink, sizeof(float) * 64));
    const int rounds = 2000;
    k_pingpong<MODE><<<2, 512>>>(20, cyc, sink);
    k_pingpong<MODE><<<2, 512>>>(rounds, cyc, sink);
    CK(cudaDeviceSynchronize());
    long long h[2];
    CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
    printf("%-34s cycles per hand-off %.1f (rank0) %.1f (rank1)\n", name, (double)h[0] / rounds, (double)h[1] / rounds);
    cudaFree(cyc); cudaFree(sink);
    return 0;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s, %d SMs\n", p.name, p.multiProcessorCount);
    const int sms = p.multiProcessorCount;
    if (run_tput<0>("scalar FFMA (16 chains)", sms, 1)) return 1;
    if (run_tput<1>("FFMA2 (8 chains)", sms, 2)) return 1;
    if (run_tput<2>("FADD2/FMUL2 alternating", sms, 2)) return 1;
    if (run_tput<3>("FFMA2 + MUFU (9 per 48)", sms, 2)) return 1;
    if (run_tput<4>("FFMA + MUFU (9 per 96)", sms, 1)) return 1;
    if (run_pp<0>("st.shared::cluster + 32 arrives")) return 1;
    if (run_pp<1>("st.async complete_tx")) return 1;
    if (run_pp<2>("barrier.cluster arrive+wait")) return 1;
    return 0;
}
