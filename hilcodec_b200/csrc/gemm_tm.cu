// "Time-major" tensor-core kernel for the narrow, high-rate layers (Cout <= 192):
//
//   D[t][co] = sum_k pre(X[b][k][t]) * W[co][k]         (tcgen05 kind::tf32, 3xTF32, fp32 in TMEM)
//
// Same math as gemm_tc.cu with the operand roles swapped: the 128 time steps of a tile are the
// M dimension (TMEM lanes) and the output channels are N.  Why a second mapping:
//   * the activation operand goes registers -> TMEM (tcgen05.st) and is read by the MMA from
//     TMEM (the "TS" form), so neither B_lo nor the activation tile is ever read from shared
//     memory by the tensor core: an SS-mode 128x128x8 tf32 MMA reads 8 KB of smem per 64 cycles,
//     which saturates the 128 B/clk port and capped gemm_tc.cu at ~50 % tensor-pipe activity;
//   * N = Cout exactly (64 / 96 / 128 / 192), no padding of narrow layers to 128 rows;
//   * the weights are the smem operand and stay resident for the whole kernel when they fit;
//   * in the epilogue a warp's 32 lanes are 32 consecutive time steps of one channel, so
//     results go straight from registers to global memory with coalesced 128-byte stores
//     (no smem staging), and the residual skip is read the same way.
// Warp roles (512 threads): 0 TMA producer (X), 1 MMA issuer, 2 TMEM alloc + weight loader,
// 4-7 epilogue, 8-11 / 12-15 two transform groups (even / odd k-blocks), each owning one of the
// two TMEM A stages.
#include "tc_ptx.cuh"

namespace hil {
namespace tm {

using namespace tc;

constexpr int BT = 128;            // time steps per tile (TMEM lanes)
constexpr int BK = 32;
constexpr int XTILE_BYTES = BK * BT * 4;   // 16 KB raw activation k-block
constexpr int NUM_THREADS = 512;
constexpr int TMEM_COLS = 512;
constexpr int A_COLS = 64;          // per A stage: 32 columns hi + 32 columns lo
constexpr int ACC_BASE = 2 * A_COLS;
constexpr int MAX_XS = 8;
constexpr int SMEM_LIMIT = 227 * 1024;

struct Params {
    int Cout, N, K, T, B;           // N = round_up(Cout, 32)
    int nkb, tiles_t;
    long long total_tiles;
    int pre;
    float pre_scale;
    const float* bias;              // [Cout] or null
    const float* R;                 // residual, same layout as Y (may alias Y), or null
    float* Y;
    long long y_bs;
    int y_rs;
    int xs;                         // X ring depth
    int w_resident;                 // 1: all k-blocks of W live in smem; 0: 2-slot ring
    int acc_stages;                 // 2 when 4*N + 128 <= 512
    uint32_t idesc;
};

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
          "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
          "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// A from tensor memory, B from a shared-memory descriptor
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tm_kernel(const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
               const __grid_constant__ CUtensorMap map_x, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t w_slot_bytes = 2u * p.N * 128u;                 // one k-block of W: hi rows then lo rows
    const int w_slots = p.w_resident ? p.nkb : 2;
    const uint32_t w_base = base;
    const uint32_t x_base = w_base + w_slots * w_slot_bytes;       // N multiple of 32 -> 1024-aligned
    const uint32_t bars = x_base + p.xs * XTILE_BYTES;
    auto xfull_bar = [&](int s) { return bars + 8u * s; };
    auto xempty_bar = [&](int s) { return bars + 8u * (MAX_XS + s); };
    auto afull_bar = [&](int g) { return bars + 8u * (2 * MAX_XS + g); };
    auto aempty_bar = [&](int g) { return bars + 8u * (2 * MAX_XS + 2 + g); };
    auto wfull_bar = [&](int s) { return bars + 8u * (2 * MAX_XS + 4 + s); };    // slot 0 doubles as "W resident"
    auto wempty_bar = [&](int s) { return bars + 8u * (2 * MAX_XS + 6 + s); };
    auto tfull_bar = [&](int a) { return bars + 8u * (2 * MAX_XS + 8 + a); };
    auto tempty_bar = [&](int a) { return bars + 8u * (2 * MAX_XS + 10 + a); };
    const uint32_t tmem_slot = bars + 8u * (2 * MAX_XS + 12);
    uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb = p.nkb;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_w_hi);
        prefetch_tmap(&map_w_lo);
        prefetch_tmap(&map_x);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < MAX_XS; ++s) { mbar_init(xfull_bar(s), 1); mbar_init(xempty_bar(s), 128); }
        for (int g = 0; g < 2; ++g) {
            mbar_init(afull_bar(g), 128);
            mbar_init(aempty_bar(g), 1);
            mbar_init(wfull_bar(g), 1);
            mbar_init(wempty_bar(g), 1);
            mbar_init(tfull_bar(g), 1);
            mbar_init(tempty_bar(g), 128);
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        const uint32_t ncols = TMEM_COLS;
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(ncols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - base));

    if (warp == 0) {
        // ===================================================================== TMA producer: activations
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const int tt = (int)(tile % p.tiles_t);
                const int b = (int)(tile / p.tiles_t);
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait<32>(xempty_bar(s), ph ^ 1);
                    mbar_arrive_expect_tx(xfull_bar(s), XTILE_BYTES);
                    tma_load_3d(&map_x, x_base + s * XTILE_BYTES, xfull_bar(s), tt * BT, kb * BK, b);
                    if (++s == p.xs) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 2) {
        // ===================================================================== weight loader
        if (lane == 0) {
            if (p.w_resident) {
                mbar_arrive_expect_tx(wfull_bar(0), (uint32_t)nkb * w_slot_bytes);
                for (int kb = 0; kb < nkb; ++kb) {
                    tma_load_2d(&map_w_hi, w_base + kb * w_slot_bytes, wfull_bar(0), kb * BK, 0);
                    tma_load_2d(&map_w_lo, w_base + kb * w_slot_bytes + p.N * 128, wfull_bar(0), kb * BK, 0);
                }
            } else {
                int s = 0;
                uint32_t ph = 0;
                for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                    for (int kb = 0; kb < nkb; ++kb) {
                        mbar_wait<32>(wempty_bar(s), ph ^ 1);
                        mbar_arrive_expect_tx(wfull_bar(s), w_slot_bytes);
                        tma_load_2d(&map_w_hi, w_base + s * w_slot_bytes, wfull_bar(s), kb * BK, 0);
                        tma_load_2d(&map_w_lo, w_base + s * w_slot_bytes + p.N * 128, wfull_bar(s), kb * BK, 0);
                        if (++s == 2) { s = 0; ph ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer
        uint32_t kc = 0;                 // running k-block counter: A stage = kc & 1, phase = (kc >> 1) & 1
        int ws = 0;
        uint32_t wph = 0;
        long long it = 0;
        if (p.w_resident) mbar_wait(wfull_bar(0), 0);
        for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
            const int acc = p.acc_stages == 2 ? (int)(it & 1) : 0;
            const uint32_t acc_ph = p.acc_stages == 2 ? (uint32_t)((it >> 1) & 1) : (uint32_t)(it & 1);
            mbar_wait(tempty_bar(acc), acc_ph ^ 1);
            tc_fence_after();
            const uint32_t d_big = tmem_base + ACC_BASE + acc * 2 * p.N;
            const uint32_t d_small = d_big + p.N;
            for (int kb = 0; kb < nkb; ++kb, ++kc) {
                const int g = kc & 1;
                mbar_wait(afull_bar(g), (kc >> 1) & 1);
                uint32_t sw_addr;
                if (p.w_resident) {
                    sw_addr = w_base + kb * w_slot_bytes;
                } else {
                    mbar_wait(wfull_bar(ws), wph);
                    sw_addr = w_base + ws * w_slot_bytes;
                }
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t a_hi = tmem_base + g * A_COLS;
                    const uint32_t a_lo = a_hi + 32;
#pragma unroll
                    for (int j = 0; j < BK / 8; ++j) {
                        // W tile: N rows of 128 bytes (32 k), 128B swizzle, 8-row groups 1024 B apart
                        const uint64_t w_hi = make_desc(sw_addr + j * 32, 16, 1024, 2);
                        const uint64_t w_lo = make_desc(sw_addr + p.N * 128 + j * 32, 16, 1024, 2);
                        umma_tf32_ts(d_big, a_hi + j * 8, w_hi, p.idesc, (kb | j) != 0);
                        umma_tf32_ts(d_small, a_lo + j * 8, w_hi, p.idesc, (kb | j) != 0);
                        umma_tf32_ts(d_small, a_hi + j * 8, w_lo, p.idesc, 1);
                    }
                    umma_commit(aempty_bar(g));
                    if (!p.w_resident) umma_commit(wempty_bar(ws));
                    if (kb == nkb - 1) umma_commit(tfull_bar(acc));
                }
                __syncwarp();
                if (!p.w_resident) {
                    if (++ws == 2) { ws = 0; wph ^= 1; }
                }
            }
        }
    } else if (warp >= 8) {
        // ===================================================================== transform: smem -> regs -> TMEM (A_hi | A_lo)
        const int g = (warp - 8) >> 2;        // k-block parity this group handles = its TMEM A stage
        const int q = warp & 3;               // TMEM lane quarter
        const int tl = q * 32 + lane;         // time step inside the tile
        const uint32_t a_hi = tmem_base + ((uint32_t)(q * 32) << 16) + g * A_COLS;
        uint32_t kc = 0;
        int xs = 0;
        uint32_t xph = 0;
        for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            for (int kb = 0; kb < nkb; ++kb, ++kc) {
                if ((int)(kc & 1) == g) {
                    mbar_wait(xfull_bar(xs), xph);
                    const float* src = reinterpret_cast<const float*>(gen_base + (x_base - base) + xs * XTILE_BYTES) + tl;
                    uint32_t hi[32], lo[32];
#pragma unroll
                    for (int k = 0; k < 32; ++k) {
                        float v = src[k * BT];   // lanes read consecutive words: conflict free
                        if (p.pre != PRE_NONE) v = elu_fast(v * p.pre_scale);
                        hi[k] = __float_as_uint(v);                         // the tensor core truncates to tf32 itself
                        lo[k] = __float_as_uint(tf32_rna(v - tf32_trunc(v)));
                    }
                    mbar_arrive(xempty_bar(xs));                            // raw tile is in registers
                    mbar_wait(aempty_bar(g), ((kc >> 1) & 1) ^ 1);          // MMAs of k-block kc-2 are done with this stage
                    tc_fence_after();
                    tmem_st32(a_hi, hi);
                    tmem_st32(a_hi + 32, lo);
                    tmem_st_wait();
                    tc_fence_before();
                    mbar_arrive(afull_bar(g));
                }
                if (++xs == p.xs) { xs = 0; xph ^= 1; }
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ===================================================================== epilogue: TMEM -> registers -> global
        const int q = warp - 4;
        long long it = 0;
        for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
            const int tt = (int)(tile % p.tiles_t);
            const int b = (int)(tile / p.tiles_t);
            const int acc = p.acc_stages == 2 ? (int)(it & 1) : 0;
            const uint32_t acc_ph = p.acc_stages == 2 ? (uint32_t)((it >> 1) & 1) : (uint32_t)(it & 1);
            const int t = tt * BT + q * 32 + lane;
            const bool t_ok = t < p.T;
            const long long off0 = (long long)b * p.y_bs + t;
            mbar_wait<64>(tfull_bar(acc), acc_ph);
            tc_fence_after();
            const uint32_t t_big = tmem_base + ((uint32_t)(q * 32) << 16) + ACC_BASE + acc * 2 * p.N;
            const int n_chunks = p.N / 32;
#pragma unroll 1
            for (int c = 0; c < n_chunks; ++c) {
                uint32_t rb[32], rs[32];
                tmem_ld32(t_big + c * 32, rb);
                tmem_ld32(t_big + p.N + c * 32, rs);
                // bias of the chunk's 32 channels: one coalesced load, broadcast per channel by shuffle
                // (a load per element sat on the critical path: 32 dependent L2 latencies per chunk)
                const int co_l = c * 32 + lane;
                const float bias_l = (p.bias && co_l < p.Cout) ? __ldg(p.bias + co_l) : 0.f;
                float* yp = p.Y + off0 + (long long)(c * 32) * p.y_rs;
                float res[32];
                if (p.R && t_ok) {   // residual loads overlap the TMEM read
                    const float* rp = p.R + off0 + (long long)(c * 32) * p.y_rs;
#pragma unroll
                    for (int j = 0; j < 32; ++j) res[j] = (c * 32 + j < p.Cout) ? rp[(long long)j * p.y_rs] : 0.f;
                }
                tmem_ld_wait();
                if (c == n_chunks - 1) {
                    tc_fence_before();
                    mbar_arrive(tempty_bar(acc));
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float bj = __shfl_sync(0xffffffffu, bias_l, j);
                    float o = (__uint_as_float(rb[j]) + __uint_as_float(rs[j])) + bj;
                    if (p.R && t_ok) o += res[j];
                    if (t_ok && c * 32 + j < p.Cout) yp[(long long)j * p.y_rs] = o;   // 32 lanes = 32 consecutive t: one 128-byte line
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        const uint32_t ncols = TMEM_COLS;
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols));
    }
}

}  // namespace tm

bool gemm_tm_usable(const PackedMat& W, const float* X, long long x_bs, int x_rs, int T, const float* Y, long long y_bs,
                    int y_rs) {
    if (!W.A_hi || !W.A_lo) return false;
    if (W.M > 192 || T < 64) return false;
    if ((x_rs & 3) || (x_bs & 3)) return false;
    if (reinterpret_cast<uintptr_t>(X) & 15) return false;
    (void)Y; (void)y_bs; (void)y_rs;
    return true;
}

cudaError_t launch_gemm_tm(const PackedMat& W, const float* X, long long x_bs, int x_rs, int B, int T, int pre,
                           float pre_scale, const float* bias, const float* R, float* Y, long long y_bs, int y_rs,
                           cudaStream_t st) {
    using namespace tm;
    if (B == 0 || T == 0) return cudaSuccess;
    Params p{};
    p.Cout = W.M; p.N = round_up(W.M, 32); p.K = W.K; p.T = T; p.B = B;
    p.nkb = (W.K + BK - 1) / BK;
    p.tiles_t = (T + BT - 1) / BT;
    p.total_tiles = (long long)p.tiles_t * B;
    p.pre = pre; p.pre_scale = (pre == PRE_SCALE_ELU) ? pre_scale : 1.0f;
    p.bias = bias; p.R = R; p.Y = Y; p.y_bs = y_bs; p.y_rs = y_rs;
    p.acc_stages = (4 * p.N + ACC_BASE <= TMEM_COLS) ? 2 : 1;
    p.idesc = make_idesc(BT, p.N, 0);
    const size_t w_slot = (size_t)2 * p.N * 128;
    const size_t fixed = 1024 + 256;
    p.w_resident = (fixed + p.nkb * w_slot + 3 * XTILE_BYTES <= (size_t)SMEM_LIMIT) ? 1 : 0;
    const size_t w_bytes = (p.w_resident ? p.nkb : 2) * w_slot;
    int xs = (int)((SMEM_LIMIT - fixed - w_bytes) / XTILE_BYTES);
    if (xs > MAX_XS) xs = MAX_XS;
    if (xs < 2) return cudaErrorInvalidValue;
    p.xs = xs;
    const size_t smem = fixed + w_bytes + (size_t)xs * XTILE_BYTES;
    static size_t attr_smem = 0;
    if (smem > attr_smem) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
        if (e != cudaSuccess) return e;
        attr_smem = SMEM_LIMIT;
    }
    CUtensorMap map_hi, map_lo, map_x;
    {
        const cuuint64_t dims[2] = {(cuuint64_t)W.Kp32, (cuuint64_t)W.Mp128};
        const cuuint64_t strides[1] = {(cuuint64_t)W.Kp32 * 4};
        const cuuint32_t box[2] = {BK, (cuuint32_t)p.N};
        if (!make_map(&map_hi, W.A_hi, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B) ||
            !make_map(&map_lo, W.A_lo, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))
            return cudaErrorInvalidValue;
    }
    {
        const cuuint64_t dims[3] = {(cuuint64_t)T, (cuuint64_t)W.K, (cuuint64_t)B};
        const cuuint64_t strides[2] = {(cuuint64_t)x_rs * 4, (cuuint64_t)x_bs * 4};
        const cuuint32_t box[3] = {BT, BK, 1};
        if (!make_map(&map_x, X, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return cudaErrorInvalidValue;
    }
    const int num_sms = device_sm_count();
    const unsigned grid = (unsigned)(p.total_tiles < num_sms ? p.total_tiles : num_sms);
    gemm_tm_kernel<<<grid, NUM_THREADS, smem, st>>>(map_hi, map_lo, map_x, p);
    return cudaGetLastError();
}

}  // namespace hil
