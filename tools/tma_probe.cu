// Standalone probe for the TMA patch staging used by the match kernels (debug aid).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <stdint.h>

#define CKC(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

#ifndef PWIDTH
#define PWIDTH 112
#endif
constexpr int PW = PWIDTH, PR = 49, COPY = PW * PR * 2, STRIDE = (COPY + 127) / 128 * 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int VARIANT>
__global__ void probe(const __grid_constant__ CUtensorMap map, int x, int y, uint16_t* out) {
    extern __shared__ __align__(128) unsigned char sm[];
    uint32_t bar = smem_u32(sm + 2 * STRIDE);
    uint32_t dst = smem_u32(sm);
    int lane = threadIdx.x & 31;
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1) : "memory");
        if (VARIANT == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (VARIANT == 1) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    if (VARIANT == 3) return;
    if (lane == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(2 * COPY) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(dst), "l"(&map), "r"(x), "r"(y), "r"(bar) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(dst + STRIDE), "l"(&map), "r"(x + 1), "r"(y), "r"(bar) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nWAIT_LOOP:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra WAIT_DONE;\nbra WAIT_LOOP;\nWAIT_DONE:\n}\n" ::"r"(bar), "r"(0) : "memory");
    const uint16_t* A = (const uint16_t*)sm;
    const uint16_t* B = (const uint16_t*)(sm + STRIDE);
    for (int i = lane; i < PW * PR; i += 32) { out[i] = A[i]; out[PW * PR + i] = B[i]; }
}

__global__ void probe_gptr(const CUtensorMap* map, int x, int y, uint16_t* out) {
    extern __shared__ __align__(128) unsigned char sm[];
    uint32_t bar = smem_u32(sm + 2 * STRIDE);
    uint32_t dst = smem_u32(sm);
    int lane = threadIdx.x & 31;
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1) : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    if (lane == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(2 * COPY) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(dst), "l"(map), "r"(x), "r"(y), "r"(bar) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(dst + STRIDE), "l"(map), "r"(x + 1), "r"(y), "r"(bar) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nWAIT_LOOP:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra WAIT_DONE;\nbra WAIT_LOOP;\nWAIT_DONE:\n}\n" ::"r"(bar), "r"(0) : "memory");
    const uint16_t* A = (const uint16_t*)sm;
    const uint16_t* B = (const uint16_t*)(sm + STRIDE);
    for (int i = lane; i < PW * PR; i += 32) { out[i] = A[i]; out[PW * PR + i] = B[i]; }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int V>
int run(const CUtensorMap& map, uint16_t* d_out, const std::vector<uint16_t>& h, int W, int pitch, int rows, int x, int y) {
    cudaFuncSetAttribute(probe<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * STRIDE + 64);
    probe<V><<<1, 32, 2 * STRIDE + 64>>>(map, x, y, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("variant %d x=%d y=%d: %s\n", V, x, y, cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    if (V == 3) return 0;
    std::vector<uint16_t> o(2 * PW * PR);
    cudaMemcpy(o.data(), d_out, o.size() * 2, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r = 0; r < PR; ++r)
        for (int k = 0; k < PW; ++k)
            for (int c = 0; c < 2; ++c) {
                int gy = y + r, gx = x + k + c;
                uint16_t exp = (gy >= 0 && gy < rows && gx >= 0 && gx < W) ? h[(size_t)gy * pitch + gx] : 0;
                if (o[c * PW * PR + r * PW + k] != exp) ++bad;
            }
    printf("   mismatches: %d\n", bad);
    return bad != 0;
}

int g_argc; char** g_argv;
int real_main();
int main(int argc, char** argv) { g_argc = argc; g_argv = argv; return real_main(); }
int real_main() {
    const int W = 1241, pitch = 1280, rows = 376 * 4;
    std::vector<uint16_t> h((size_t)pitch * rows);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (uint16_t)((i * 2654435761u) >> 17);
    uint16_t *d, *d_out;
    CKC(cudaMalloc(&d, h.size() * 2));
    CKC(cudaMalloc(&d_out, 2 * PW * PR * 2));
    CKC(cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    CKC(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    printf("entry point %p status %d\n", p, (int)q);
    CUtensorMap map;
    const cuuint64_t gdim[2] = {(cuuint64_t)W, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)pitch * 2};
    const cuuint32_t box[2] = {PW, PR};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = ((EncodeTiledFn)p)(&map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, d, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode -> %d\n", (int)r);
    extern int g_argc; extern char** g_argv;
    int x = g_argc > 1 ? atoi(g_argv[1]) : 100, y = g_argc > 2 ? atoi(g_argv[2]) : 50;
    int variant = g_argc > 3 ? atoi(g_argv[3]) : 1;
    if (variant == 4) {
        CUtensorMap* dmap;
        CKC(cudaMalloc(&dmap, sizeof(CUtensorMap)));
        CKC(cudaMemcpy(dmap, &map, sizeof(CUtensorMap), cudaMemcpyHostToDevice));
        cudaFuncSetAttribute(probe_gptr, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * STRIDE + 64);
        probe_gptr<<<1, 32, 2 * STRIDE + 64>>>(dmap, x, y, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        printf("variant gptr x=%d y=%d: %s\n", x, y, cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        std::vector<uint16_t> o(2 * PW * PR);
        cudaMemcpy(o.data(), d_out, o.size() * 2, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int r = 0; r < PR; ++r) for (int k = 0; k < PW; ++k) for (int c = 0; c < 2; ++c) {
            int gy = y + r, gx = x + k + c;
            uint16_t exp = (gy >= 0 && gy < rows && gx >= 0 && gx < W) ? h[(size_t)gy * pitch + gx] : 0;
            if (o[c * PW * PR + r * PW + k] != exp) ++bad;
        }
        printf("   mismatches: %d\n", bad);
        return bad != 0;
    }
    return run<1>(map, d_out, h, W, pitch, rows, x, y);
}
