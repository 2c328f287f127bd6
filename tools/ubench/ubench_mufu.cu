// MUFU throughput on sm_100a: sqrt.approx vs rsqrt.approx vs rcp.approx vs ex2.approx, warp instructions per clock per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_mufu ubench_mufu.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void __launch_bounds__(512, 1) k(float *out, int n, float seed) {
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x * 1e-3f + i;
    for (int it = 0; it < n; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (OP == 1) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (OP == 3) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (OP == 4) {  // sqrt as x * rsqrt(max(x, tiny)): one MUFU.RSQ + FMNMX + FMUL
                float t;
                asm volatile("max.f32 %0, %1, 0f0DA24260;" : "=f"(t) : "f"(a[i]));
                asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(t));
                asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(t));
            }
            if (OP == 5) asm volatile("sqrt.approx.f32 %0, %0;" : "+f"(a[i]));  // with denormal scaling
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
void run(const char *name, float *out, int sms) {
    const int n = 4096;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<OP><<<sms, 512>>>(out, 64, 1.5f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<OP><<<sms, 512>>>(out, n, 1.5f);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double warp_instr_per_sm = (double)n * 8 * 16;  // 16 warps per SM
    printf("%-28s %8.3f ms  %.1f ns per warp-op per SM  (%.2f clk at %d MHz nominal)\n", name, ms,
           ms * 1e6 / warp_instr_per_sm, ms * 1e-3 * clk_khz * 1e3 / warp_instr_per_sm, clk_khz / 1000);
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float *out;
    cudaMalloc(&out, sizeof(float) * sms * 512);
    run<0>("sqrt.approx.ftz", out, sms);
    run<5>("sqrt.approx (no ftz)", out, sms);
    run<1>("rsqrt.approx.ftz", out, sms);
    run<2>("rcp.approx.ftz", out, sms);
    run<3>("ex2.approx.ftz", out, sms);
    run<4>("x*rsqrt(max(x,tiny))", out, sms);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
