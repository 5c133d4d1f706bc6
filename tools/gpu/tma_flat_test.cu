// Checks the two tensor maps of gemm_h.cu's flat tiles in isolation: load of a {t, clip, k} box (strides not
// monotonic) and store of a {t, clip, m} box with SWIZZLE_128B and a 32-byte inner extent.  Build: see build_peaks.sh.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../hilcodec_b200/csrc/tc_ptx.cuh"
using namespace hil::tc;

__global__ void load_kernel(const __grid_constant__ CUtensorMap map, float* out, int clip0, int k0) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    const uint32_t bar = base + 16384;
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(bar, 16384);
        tma_load_3d(&map, base, bar, 0, clip0, k0);
    }
    mbar_wait(bar, 0);
    const float* s = reinterpret_cast<const float*>(smem + (base - smem_u32(smem)));
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) out[i] = s[i];
}

__global__ void store_kernel(const __grid_constant__ CUtensorMap map, int clip0, int m0, int swz) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    float* s = reinterpret_cast<float*>(smem + (base - smem_u32(smem)));
    // value of (row m, column c) = m * 100 + c, written at the swizzled position the epilogue uses
    for (int i = threadIdx.x; i < 128 * 32; i += blockDim.x) {
        const int m = i / 32, c = i % 32;
        // 16-byte chunk index inside the 128-byte row of m: 128B swizzle XORs it with (m & 7); 32B swizzle XORs the chunk's
        // low bit with address bit 7 = (m & 1)
        const int chunk = c / 4, sw = swz == 1 ? (m & 7) : swz == 2 ? (m & 1) : 0;
        s[m * 32 + ((chunk ^ sw) * 4) + (c & 3)] = (float)(m * 100 + c);
    }
    fence_proxy_async();
    __syncthreads();
    if (threadIdx.x == 0) {
        tma_store_3d(&map, base, 0, clip0, m0);
        tma_commit();
        tma_wait_all();
    }
}

int main(int argc, char** argv) {
    const int mode = argc > 1 ? atoi(argv[1]) : 0;   // output map: 0 SWIZZLE_NONE, 1 SWIZZLE_128B, 2 SWIZZLE_32B
    const int T = 8, B = 40, K = 96, M = 256;
    std::vector<float> hx((size_t)B * K * T);
    for (int b = 0; b < B; ++b) for (int k = 0; k < K; ++k) for (int t = 0; t < T; ++t) hx[((size_t)b * K + k) * T + t] = b * 10000 + k * 10 + t;
    float *dx, *dout, *dy;
    cudaMalloc(&dx, hx.size() * 4); cudaMemcpy(dx, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice);
    cudaMalloc(&dout, 4096 * 4);
    cudaMalloc(&dy, (size_t)B * M * T * 4); cudaMemset(dy, 0, (size_t)B * M * T * 4);
    CUtensorMap mx, my;
    {
        const cuuint64_t dims[3] = {(cuuint64_t)T, (cuuint64_t)B, (cuuint64_t)K};
        const cuuint64_t strides[2] = {(cuuint64_t)K * T * 4, (cuuint64_t)T * 4};
        const cuuint32_t box[3] = {(cuuint32_t)T, 16, 32};
        printf("x map ok=%d\n", (int)make_map(&mx, dx, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE));
    }
    cudaFuncSetAttribute(load_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 20000);
    cudaFuncSetAttribute(store_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 20000);
    load_kernel<<<1, 128, 20000>>>(mx, dout, 32, 64);   // clips 32 .. 47 (40 .. 47 out of bounds), k 64 .. 95
    printf("load: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    std::vector<float> ho(4096);
    cudaMemcpy(ho.data(), dout, 4096 * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int k = 0; k < 32; ++k) for (int c = 0; c < 16; ++c) for (int t = 0; t < T; ++t) {
        const int b = 32 + c;
        const float want = b < B ? b * 10000 + (64 + k) * 10 + t : 0.f;
        if (ho[(k * 16 + c) * 8 + t] != want) { if (bad < 5) printf("  load mismatch k=%d c=%d t=%d got %g want %g\n", k, c, t, ho[(k * 16 + c) * 8 + t], want); ++bad; }
    }
    printf("load mismatches: %d\n", bad);
    {
        const int swz = mode;
        const cuuint64_t dims[3] = {(cuuint64_t)T, (cuuint64_t)B, (cuuint64_t)M};
        const cuuint64_t strides[2] = {(cuuint64_t)M * T * 4, (cuuint64_t)T * 4};
        const cuuint32_t box[3] = {(cuuint32_t)T, 4, 128};
        printf("y map (swizzle %d) ok=%d\n", swz, (int)make_map(&my, dy, 3, dims, strides, box, swz == 1 ? CU_TENSOR_MAP_SWIZZLE_128B : swz == 2 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE));
        cudaMemset(dy, 0, (size_t)B * M * T * 4);
        store_kernel<<<1, 128, 20000>>>(my, 36, 128, swz);   // clips 36 .. 39, rows 128 .. 255
        printf("store: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
        std::vector<float> hy((size_t)B * M * T);
        cudaMemcpy(hy.data(), dy, hy.size() * 4, cudaMemcpyDeviceToHost);
        bad = 0;
        for (int b = 0; b < B; ++b) for (int m = 0; m < M; ++m) for (int t = 0; t < T; ++t) {
            float want = 0.f;
            if (b >= 36 && b < 40 && m >= 128) want = (m - 128) * 100 + (b - 36) * 8 + t;
            const float got = hy[((size_t)b * M + m) * T + t];
            if (got != want) { if (bad < 5) printf("  store mismatch b=%d m=%d t=%d got %g want %g\n", b, m, t, got, want); ++bad; }
        }
        printf("store mismatches (swizzle %d): %d\n", swz, bad);
    }
    return 0;
}
