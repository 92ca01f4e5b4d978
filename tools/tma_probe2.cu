// Probe 2: same load through libcu++'s documented wrappers (CUDA programming guide example).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <cstdio>
#include <vector>
#include <stdint.h>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
#ifndef PWIDTH
#define PWIDTH 112
#endif
constexpr int PW = PWIDTH, PR = 49;

__global__ void probe(const __grid_constant__ CUtensorMap map, int x, int y, uint16_t* out) {
    __shared__ alignas(128) uint16_t buf[PR][PW];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        cde::cp_async_bulk_tensor_2d_global_to_shared(&buf, &map, x, y, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(buf));
    } else token = bar.arrive();
    bar.wait(std::move(token));
    for (int i = threadIdx.x; i < PW * PR; i += blockDim.x) out[i] = (&buf[0][0])[i];
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv) {
    const int W = 1241, pitch = 1280, rows = 376 * 4;
    int x = argc > 1 ? atoi(argv[1]) : 100, y = argc > 2 ? atoi(argv[2]) : 50;
    std::vector<uint16_t> h((size_t)pitch * rows);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (uint16_t)((i * 2654435761u) >> 17);
    uint16_t *d, *d_out;
    cudaMalloc(&d, h.size() * 2); cudaMalloc(&d_out, PW * PR * 2);
    cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    CUtensorMap map;
    const cuuint64_t gdim[2] = {(cuuint64_t)W, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)pitch * 2};
    const cuuint32_t box[2] = {PW, PR};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = ((EncodeTiledFn)p)(&map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, d, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode -> %d\n", (int)r);
    probe<<<1, 128>>>(map, x, y, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("libcu++ probe box %d x=%d y=%d: %s\n", PW, x, y, cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<uint16_t> o(PW * PR);
    cudaMemcpy(o.data(), d_out, o.size() * 2, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int rr = 0; rr < PR; ++rr) for (int k = 0; k < PW; ++k) {
        int gy = y + rr, gx = x + k;
        uint16_t exp = (gy >= 0 && gy < rows && gx >= 0 && gx < W) ? h[(size_t)gy * pitch + gx] : 0;
        if (o[rr * PW + k] != exp) ++bad;
    }
    printf("   mismatches: %d\n", bad);
    return bad != 0;
}
