// Minimal CPU emulation of the CUDA constructs used by the small FP32 kernels of hilcodec_b200 (gemm_skinny.cu,
// rvq.cu's per-stage kernel), so that their SOURCE TEXT (extracted by tests/test_kernel_emulation.py) can be compiled
// with g++ and run here, where there is no GPU: one std::thread per CUDA thread, std::barrier for __syncthreads /
// __syncwarp, an exchange buffer for warp shuffles.  Test infrastructure only.  It checks indexing, reduction order
// and arithmetic, not memory-model subtleties (every access is sequentially consistent here).
#pragma once
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#include <cstddef>
using std::size_t;

struct uint3e { unsigned x = 0, y = 0, z = 0; };
inline thread_local uint3e threadIdx, blockIdx;
inline uint3e blockDim, gridDim;

struct alignas(8) float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
template <class T> inline T __ldg(const T* p) { return *p; }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline float fmaxf(float a, float b) { return std::fmax(a, b); }
inline float fminf(float a, float b) { return std::fmin(a, b); }
// compiled with -ffp-contract=off: these stay separately rounded operations
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fdiv_rn(float a, float b) { return a / b; }

alignas(16) inline float g_dyn_smem[64 * 1024];  // 256 KB: "dynamic shared memory" of the one CTA that is running

namespace emu {
inline std::barrier<>* cta_bar = nullptr;
inline std::vector<std::unique_ptr<std::barrier<>>> warp_bar;
inline uint32_t xchg[64][32];
}  // namespace emu

inline void __syncthreads() { emu::cta_bar->arrive_and_wait(); }
inline void __syncwarp() { emu::warp_bar[threadIdx.x >> 5]->arrive_and_wait(); }
template <class T> inline T __shfl_xor_sync(unsigned, T v, int o) {
    static_assert(sizeof(T) == 4, "32-bit shuffles only");
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    std::memcpy(&emu::xchg[w][l], &v, 4);
    emu::warp_bar[w]->arrive_and_wait();
    T r;
    std::memcpy(&r, &emu::xchg[w][l ^ o], 4);
    emu::warp_bar[w]->arrive_and_wait();
    return r;
}

struct dim3 {
    unsigned x = 1, y = 1, z = 1;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

// run `kernel()` for every CTA of the grid (CTAs one after the other, the threads of a CTA concurrently); `rows` > 1 makes
// the block two-dimensional: blockDim = (threads, rows), thread t of the CTA is (t % threads, t / threads)
template <class F> void emu_launch(unsigned gx, unsigned gy, unsigned threads, F kernel, unsigned gz = 1, unsigned rows = 1) {
    gridDim = {gx, gy, gz};
    blockDim = {threads, rows, 1};
    const unsigned tx = threads;
    threads *= rows;
    for (unsigned bz = 0; bz < gz; ++bz)
    for (unsigned by = 0; by < gy; ++by)
        for (unsigned bx = 0; bx < gx; ++bx) {
            std::barrier<> cta((std::ptrdiff_t)threads);
            emu::cta_bar = &cta;
            emu::warp_bar.clear();
            for (unsigned w = 0; w < (threads + 31) / 32; ++w)
                emu::warp_bar.push_back(std::make_unique<std::barrier<>>((std::ptrdiff_t)(threads - 32 * w < 32 ? threads - 32 * w : 32)));
            std::vector<std::thread> ts;
            for (unsigned t = 0; t < threads; ++t)
                ts.emplace_back([=, &kernel] {
                    threadIdx = {t % tx, t / tx, 0};
                    blockIdx = {bx, by, bz};
                    kernel();
                    emu::cta_bar->arrive_and_drop();               // a thread that returned no longer takes part
                    emu::warp_bar[t >> 5]->arrive_and_drop();
                });
            for (auto& th : ts) th.join();
        }
}
