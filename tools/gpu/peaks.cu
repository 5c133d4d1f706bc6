// Micro-benchmarks of the two compute peaks this repo's rooflines need beside MEASURED_PEAKS.json (HBM copy, cuBLAS
// bf16): the FP32 FFMA pipe (what the FP32 kernels -- RVQ search, skinny / fallback GEMMs -- run on) and the tcgen05
// tensor pipe in the kinds and instruction shapes the GEMM kernels issue (kind::f16 and kind::tf32, M = 128, N = 256 / 128,
// and the 3-MMA "fp32-accurate product" pattern of gemm_h.cu).  No data dependence on memory: operands are zeroed shared
// memory, so the numbers are issue-rate ceilings, timed with CUDA events.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/gpu/peaks tools/gpu/peaks.cu   (tools/gpu/build_peaks.sh)
//   gpurun -- 'tools/gpu/peaks > gpurun_out/peaks.json'
#include <cstdio>
#include <vector>

#include "../../hilcodec_b200/csrc/h_split.cuh"

using namespace hil;
using namespace hil::tc;
using namespace hil::th;

__global__ void __launch_bounds__(256) ffma_kernel(float* out, int iters) {
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3f + i;
    const float b = 0.999f, c = 1e-3f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], b, c);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) ffma2_kernel(float* out, int iters) {   // packed fp32 (FFMA2)
    f32x2 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = pk2(threadIdx.x * 1e-3f + i, 1.f + i);
    const f32x2 b = pk2(0.999f, 0.999f), c = pk2(1e-3f, 1e-3f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = ffma2(a[i], b, c);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { float x, y; upk2(a[i], x, y); s += x + y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mode 0: kind::f16 M128 N256;  1: kind::f16 M128 N128;  2: gemm_h.cu's pattern per k16 step (N256 + N128);
// 3: kind::tf32 M128 N256 (K = 8 per instruction);  4: kind::tf32 pattern of gemm_tc.cu / stft_tc.cu (N256 + N128)
__global__ void __launch_bounds__(128, 1) mma_kernel(int mode, int iters) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar = base + 96 * 1024, slot = bar + 16;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(gen)[i] = 0u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        const uint32_t ncols = 512;
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(ncols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    fence_proxy_async();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + (slot - base));
    if (warp == 0) {
        if (elect_one()) {   // elect.sync: no per-instruction waterfall around the tcgen05 ops (as the kernels issue them)
            const bool f16 = mode <= 2;
            // K-major A (SWIZZLE_64B for fp16 as in gemm_h.cu, SWIZZLE_128B for tf32), MN-major / K-major B as the kernels use them
            const uint64_t a = f16 ? make_desc(base, 16, 512, 4) : make_desc(base, 16, 1024, 2);
            const uint64_t b = f16 ? make_desc(base + 32 * 1024, 4096, 1024, 2) : make_desc(base + 32 * 1024, 16, 1024, 2);
            const uint32_t i256 = f16 ? make_idesc_f16(128, 256) : make_idesc(128, 256, 0);
            const uint32_t i128 = f16 ? make_idesc_f16(128, 128) : make_idesc(128, 128, 0);
            for (int it = 0; it < iters; ++it) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (mode == 0 || mode == 3) {
                        if (f16) umma_f16(tmem, a, b, i256, 1); else umma_tf32(tmem, a, b, i256, 1);
                    } else if (mode == 1) {
                        umma_f16(tmem, a, b, i128, 1);
                    } else {
                        if (f16) { umma_f16(tmem, a, b, i256, 1); umma_f16(tmem + 128, a, b, i128, 1); }
                        else { umma_tf32(tmem, a, b, i256, 1); umma_tf32(tmem + 128, a, b, i128, 1); }
                    }
                }
            }
            umma_commit(bar);
        }
        __syncwarp();
        mbar_wait<64>(bar, 0);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t ncols = 512;
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(ncols));
    }
}

template <class F>
static double time_ms(F launch, int reps) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    int sms = 0, clock_khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0);
    float* out = nullptr;
    cudaMalloc(&out, (size_t)sms * 8 * 256 * 4 * sizeof(float));
    std::printf("{\"sms\": %d, \"clock_mhz_attr\": %.0f", sms, clock_khz / 1e3);
    {
        const int iters = 4096, blocks = sms * 8;
        const double ms = time_ms([&] { ffma_kernel<<<blocks, 256>>>(out, iters); }, 5);
        std::printf(", \"fp32_ffma_tflops\": %.2f", 2.0 * blocks * 256 * (double)iters * 128 / (ms * 1e-3) / 1e12);
        const double ms2 = time_ms([&] { ffma2_kernel<<<blocks, 256>>>(out, iters); }, 5);
        std::printf(", \"fp32_ffma2_packed_tflops\": %.2f", 2.0 * blocks * 256 * (double)iters * 128 / (ms2 * 1e-3) / 1e12);
    }
    const size_t smem = 1024 + 96 * 1024 + 64;
    cudaFuncSetAttribute(mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const char* names[5] = {"tcgen05_f16_m128n256_tflops", "tcgen05_f16_m128n128_tflops", "tcgen05_f16_split_pattern_tflops",
                            "tcgen05_tf32_m128n256_tflops", "tcgen05_tf32_split_pattern_tflops"};
    for (int mode = 0; mode < 5; ++mode) {
        const int iters = 2048;
        const double ms = time_ms([&] { mma_kernel<<<sms, 128, smem>>>(mode, iters); }, 5);
        const double k = mode <= 2 ? 16.0 : 8.0;
        const double n = (mode == 0 || mode == 3) ? 256.0 : mode == 1 ? 128.0 : 384.0;
        const double flops = 2.0 * 128 * n * k * 8.0 * iters * sms;
        std::printf(", \"%s\": %.1f", names[mode], flops / (ms * 1e-3) / 1e12);
        if (mode == 2 || mode == 4)   // fp32-accurate products per second: 3 MMAs (hi*hi, hi*lo, lo*hi) per product
            std::printf(", \"%s\": %.1f", mode == 2 ? "f16_split_fp32_accurate_tflops" : "tf32_split_fp32_accurate_tflops",
                        flops / 3.0 / (ms * 1e-3) / 1e12);
    }
    cudaError_t e = cudaDeviceSynchronize();
    std::printf(", \"cuda_status\": \"%s\"}\n", cudaGetErrorString(e));
    return e == cudaSuccess ? 0 : 1;
}
