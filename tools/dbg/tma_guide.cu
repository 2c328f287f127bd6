// The CUDA programming guide's TMA example, verbatim API (libcu++ wrappers), as a control experiment.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <cstdio>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
constexpr int W = 64, H = 64, GW = 288, GH = 204;

__global__ void kernel(const __grid_constant__ CUtensorMap tensor_map, int x, int y, int *out) {
    __shared__ alignas(128) int smem_buffer[H][W];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) {
        init(&bar, blockDim.x);
        cde::fence_proxy_async_shared_cta();
    }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        cde::cp_async_bulk_tensor_2d_global_to_shared(&smem_buffer, &tensor_map, x, y, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(smem_buffer));
    } else {
        token = bar.arrive();
    }
    bar.wait(std::move(token));
    for (int i = threadIdx.x; i < H * W; i += blockDim.x) out[i] = smem_buffer[i / W][i % W];
}

// 1-D bulk copy control (no tensor map)
__global__ void kernel1d(const int *src, int *out) {
    __shared__ alignas(128) int buf[1024];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        cuda::memcpy_async(buf, src, cuda::aligned_size_t<16>(sizeof(buf)), bar);
        token = bar.arrive();
    } else token = bar.arrive();
    bar.wait(std::move(token));
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) out[i] = buf[i];
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
    int *d, *out;
    cudaMalloc(&d, GW * GH * 4); cudaMalloc(&out, W * H * 4);
    std::vector<int> h(GW * GH);
    for (int i = 0; i < GW * GH; ++i) h[i] = i;
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    kernel1d<<<1, 128>>>(d, out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("1d bulk copy: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    void *fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    CUtensorMap tm{};
    cuuint64_t size[2] = {GW, GH}; cuuint64_t stride[1] = {GW * 4};
    cuuint32_t box[2] = {W, H}; cuuint32_t es[2] = {1, 1};
    CUresult r = ((EncodeFn)fp)(&tm, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, d, size, stride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d query %d\n", (int)r, (int)q);
    const unsigned char *b = reinterpret_cast<const unsigned char *>(&tm);
    for (int i = 0; i < 128; ++i) printf("%02x%s", b[i], (i % 32 == 31) ? "\n" : "");
    kernel<<<1, 128>>>(tm, 32, 16, out);
    e = cudaDeviceSynchronize();
    printf("guide TMA 2d: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<int> o(W * H);
    cudaMemcpy(o.data(), out, o.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int j = 0; j < H; ++j) for (int i = 0; i < W; ++i) if (o[j * W + i] != (16 + j) * GW + 32 + i) ++bad;
    printf("mismatches %d\n", bad);
    return 0;
}
