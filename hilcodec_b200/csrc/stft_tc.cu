// CausalSTFT + magnitude + clamp + log on the tensor pipe (tcgen05; fp16 hi/lo splits by default, 3xTF32 with
// HILCODEC_GEMM=tf32).
//
//   S[2F][t] = sum_k Wdft[2F][k] * wav[t*hop + k],  y[f][t] = log(max(sqrt(re_f^2 + im_f^2), 1e-5))
//
// Replaces CausalSTFT.forward (causal_layers.py:135-144: the DFT as a strided Conv1d) and the
// clamp/log of SpecBlock.forward (streaming.py:351).  Same pipeline as gemm_tc.cu (TMA ->
// transform warps -> tcgen05.mma into two TMEM accumulators -> epilogue warps -> TMA store);
// what differs:
//   * A = DFT basis with rows interleaved (cos_f, sin_f), so re/im of one bin sit in adjacent
//     TMEM lanes and the magnitude is one warp shuffle in the epilogue;
//   * B = the im2col view of the waveform, K-major: row t of the tile is the n_fft-sample
//     window starting at t*hop.  For hop % 4 == 0 that is a plain (overlapping-row) tensor map
//     and TMA loads it; for hop 1 / 2 (row stride not a multiple of 16 bytes) the transform
//     warps gather the windows through L1 and write the swizzled tile themselves;
//   * the epilogue stores F rows per 128 accumulator rows (box 32 x 64).
#include <cstdlib>
#include <cstring>

#include "h_split.cuh"

namespace hil {
namespace stft {

using namespace tc;
using th::LO_SCALE;

constexpr int BM = 128, BN = 128, BK = 32;
constexpr int STAGES = 3;
constexpr int TILE_BYTES = BM * BK * 4;
constexpr int STAGE_BYTES = 4 * TILE_BYTES;      // A_hi, A_lo, B_hi, B_lo
// kH = true (fp16 hi/lo splits, same number format and MMA pattern as gemm_h.cu: twice the tf32 rate): a stage is
// A_hi | A_lo | B_hi | B_lo as K-major fp16 tiles (128 rows x 32 k = 64-byte rows, SWIZZLE_64B, 8 KB each) + the raw fp32
// window tile the TMA delivers (16 KB) = 48 KB, four stages
constexpr int H_TILE = BM * BK * 2;              // 8 KB
constexpr int H_STAGE_BYTES = 4 * H_TILE + TILE_BYTES;
constexpr int H_STAGES = 4;
constexpr uint32_t IDESC_H_N128 = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);        // f16 x f16 -> f32, A and B K-major
constexpr uint32_t IDESC_H_N256 = (1u << 4) | ((uint32_t)(2 * BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
constexpr int OUT_BYTES = 64 * 32 * 4;           // 64 bins x 32 columns
constexpr int NUM_THREADS = 512;
constexpr int NUM_XFORM = 256;
constexpr int NUM_EPI = 128;
constexpr int TMEM_COLS = 512;
constexpr int SPAN_FLOATS = 512;                 // gather mode: (BN - 1) * hop + n_fft samples of one tile (hop 1 / 2: 191 / 382)
constexpr size_t SMEM_BYTES = 1024 + (size_t)STAGES * STAGE_BYTES + 2 * OUT_BYTES + 256 + 2 * SPAN_FLOATS * 4;
static_assert((size_t)H_STAGES * H_STAGE_BYTES <= (size_t)STAGES * STAGE_BYTES, "the fp16 configuration fits the same allocation");
constexpr uint32_t IDESC = make_idesc(BM, BN, 0);  // A and B both K-major
constexpr uint32_t IDESC_N256 = make_idesc(BM, 2 * BN, 0);

struct Params {
    int M, F, K, T, B, hop;
    int num_m, tiles_t;
    long long total_tiles;
    int gather;          // transform warps build B themselves (hop % 4 != 0): 1 from a staged span, 2 from global memory
    int exact_log;       // 1: logf (HILCODEC_STFT_LOGF=1), 0: lg2.approx * ln 2
    const float* wav;    // window base (first sample of the window of frame 0)
    long long w_bs;      // batch stride of wav
    float c_big;         // kH: 2^-s of the DFT basis' fp16 scaling (gemm_h.cu)
    int flat_t;          // flat tiles (gemm_h.cu): windows per clip, 128 / flat_t whole clips per tile; T = all flat columns, B = 1
};

template <bool kH>
__global__ void __launch_bounds__(NUM_THREADS, 1)
stft_tc_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
               const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_y, const Params p) {
    constexpr int STAGES = kH ? H_STAGES : stft::STAGES;
    constexpr int STAGE_BYTES = kH ? H_STAGE_BYTES : stft::STAGE_BYTES;
    constexpr int A_BYTES = kH ? H_TILE : TILE_BYTES;            // one A plane of a stage
    constexpr int B_OFF = 2 * A_BYTES;                           // B_hi of a stage
    constexpr int RAW_OFF = kH ? 4 * H_TILE : 2 * TILE_BYTES;    // where the TMA puts the fp32 window tile (kH: behind the fp16 tiles)
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t out_base = base + stft::STAGES * stft::STAGE_BYTES;
    const uint32_t bars = out_base + 2 * OUT_BYTES;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto xform_bar = [&](int s) { return bars + 8u * (STAGES + s); };
    auto empty_bar = [&](int s) { return bars + 8u * (2 * STAGES + s); };
    auto tfull_bar = [&](int a) { return bars + 8u * (3 * STAGES + a); };
    auto tempty_bar = [&](int a) { return bars + 8u * (3 * STAGES + 2 + a); };
    const uint32_t tmem_slot = bars + 8u * (3 * STAGES + 4);
    uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
    float* span_base = reinterpret_cast<float*>(gen_base + (bars - base) + 256);   // 2 x SPAN_FLOATS (gather mode)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb = p.K / BK;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_a_hi);
        prefetch_tmap(&map_a_lo);
        prefetch_tmap(&map_x);
        prefetch_tmap(&map_y);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(xform_bar(s), NUM_XFORM / 32);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), NUM_EPI);
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
        // ===================================================================== TMA producer
        if (elect_one()) {
            int s = 0;
            uint32_t ph = 0;
            for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const int m_blk = (int)(tile % p.num_m);
                const long long rest = tile / p.num_m;
                const int tt = (int)(rest % p.tiles_t);
                const int b = (int)(rest / p.tiles_t);
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait<32>(empty_bar(s), ph ^ 1);
                    const uint32_t st = base + s * STAGE_BYTES;
                    mbar_arrive_expect_tx(full_bar(s), 2 * A_BYTES + (p.gather ? 0 : TILE_BYTES));
                    tma_load_2d(&map_a_hi, st, full_bar(s), kb * BK, m_blk * BM);
                    tma_load_2d(&map_a_lo, st + A_BYTES, full_bar(s), kb * BK, m_blk * BM);
                    if (!p.gather) {
                        if (p.flat_t) tma_load_3d(&map_x, st + RAW_OFF, full_bar(s), kb * BK, 0, tt * (BN / p.flat_t));
                        else tma_load_3d(&map_x, st + RAW_OFF, full_bar(s), kb * BK, tt * BN, b);
                    }
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer
        int s = 0;
        uint32_t ph = 0;
        long long it = 0;
        for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
            const int acc = (int)(it & 1);
            const uint32_t acc_ph = (uint32_t)((it >> 1) & 1);
            mbar_wait(tempty_bar(acc), acc_ph ^ 1);
            tc_fence_after();
            const uint32_t d_big = tmem_base + acc * 2 * BN;
            const uint32_t d_small = d_big + BN;
            for (int kb = 0; kb < nkb; ++kb) {
                const uint32_t st = base + s * STAGE_BYTES;
                // one wait per k-block on the issuer's serial path: every transform thread waited for full_bar(s) (the
                // weight tile's TMA bytes included) before it arrived on xform_bar(s), so this wait covers both
                mbar_wait(xform_bar(s), ph);
                tc_fence_after();
                if (elect_one()) {
                    // B_hi (rows 0..127) and B_lo (rows 128..255) are adjacent K-major tiles: one N = 256 MMA
                    // computes A_hi*[B_hi | B_lo] into [big | small], a second N = 128 MMA adds A_lo*B_hi.
                    if constexpr (kH) {
#pragma unroll
                        for (int j = 0; j < BK / 16; ++j) {   // 64-byte rows, SWIZZLE_64B: 8-row groups 512 B apart, +32 B per k16
                            const uint64_t a_hi = make_desc(st + j * 32, 16, 512, 4);
                            const uint64_t a_lo = make_desc(st + A_BYTES + j * 32, 16, 512, 4);
                            const uint64_t b_hl = make_desc(st + B_OFF + j * 32, 16, 512, 4);
                            th::umma_f16(d_big, a_hi, b_hl, IDESC_H_N256, (kb | j) != 0);
                            th::umma_f16(d_small, a_lo, b_hl, IDESC_H_N128, 1);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < BK / 8; ++j) {
                            const uint64_t a_hi = make_desc(st + j * 32, 16, 1024, 2);
                            const uint64_t a_lo = make_desc(st + TILE_BYTES + j * 32, 16, 1024, 2);
                            const uint64_t b_hl = make_desc(st + 2 * TILE_BYTES + j * 32, 16, 1024, 2);
                            umma_tf32(d_big, a_hi, b_hl, IDESC_N256, (kb | j) != 0);
                            umma_tf32(d_small, a_lo, b_hl, IDESC, 1);
                        }
                    }
                    umma_commit(empty_bar(s));
                    if (kb == nkb - 1) umma_commit(tfull_bar(acc));
                }
                __syncwarp();
                if (++s == STAGES) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp >= 8) {
        // ===================================================================== transform: B_lo (and B_hi when gathering)
        const int xt = threadIdx.x - 256;
        int s = 0;
        uint32_t ph = 0;
        uint32_t tcount = 0;
        for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tcount) {
            const long long rest = tile / p.num_m;
            const int tt = (int)(rest % p.tiles_t);
            const int b = (int)(rest / p.tiles_t);
            // Gather mode (hop 1 / 2: the im2col rows are not 16-byte aligned, no tensor map): the tile's 128 windows
            // overlap almost completely -- B[t][k] = wav[(t0 + t) * hop + k] -- so its whole operand comes from ONE span
            // of (BN - 1) * hop + n_fft samples.  Stage the span in shared memory with coalesced loads (double-buffered
            // by tile parity, one barrier of the 8 transform warps per tile) and build the k-blocks from it; the
            // per-k-block scalar global gathers this replaces were what the high-rate stages waited on (ncu: long
            // scoreboard on the first use of every gathered value).
            float* span = span_base + (tcount & 1u) * SPAN_FLOATS;
            if (p.gather == 1) {
                const int span_len = (BN - 1) * p.hop + p.K;
                const long long g0 = (long long)tt * BN * p.hop;                      // first sample of the tile's first window
                const long long g_end = (long long)(p.T - 1) * p.hop + p.K;            // samples the clip's windows cover
                const float* src = p.wav + (long long)b * p.w_bs + g0;
                for (int i = xt; i < span_len; i += NUM_XFORM) span[i] = (g0 + i < g_end) ? __ldg(src + i) : 0.f;
                asm volatile("bar.sync 2, 256;" ::: "memory");
            }
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(full_bar(s), ph);   // also: the slot is free (the producer waited for its MMAs)
                if constexpr (kH) {
                    // fp32 window tile (TMA, 128-byte rows, SWIZZLE_128B) or gathered values -> fp16 hi / lo tiles
                    // (64-byte rows, SWIZZLE_64B: 16-byte chunk c of row t sits at chunk c ^ ((t >> 1) & 3))
                    const float4* raw = reinterpret_cast<const float4*>(gen_base + s * STAGE_BYTES + RAW_OFF);
                    uint8_t* bhi = gen_base + s * STAGE_BYTES + B_OFF;
#pragma unroll
                    for (int i = 0; i < TILE_BYTES / 16 / NUM_XFORM; ++i) {
                        const int item = xt + i * NUM_XFORM;
                        const int t = item >> 3;
                        int k4;
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (p.gather) {
                            k4 = item & 7;
                            const int tg = tt * BN + t;
                            if (tg < p.T) {
                                if (p.gather == 1) {
                                    const float* src = span + t * p.hop + kb * BK + k4 * 4;
                                    v.x = src[0]; v.y = src[1]; v.z = src[2]; v.w = src[3];
                                } else {
                                    const float* src = p.wav + (long long)b * p.w_bs + (long long)tg * p.hop + kb * BK + k4 * 4;
                                    v.x = __ldg(src); v.y = __ldg(src + 1); v.z = __ldg(src + 2); v.w = __ldg(src + 3);
                                }
                            }
                        } else {
                            k4 = (item & 7) ^ (t & 7);   // the float4 at physical index `item` holds k = 4 * k4 .. 4 * k4 + 3 of row t
                            v = raw[item];
                        }
                        uint32_t h01, h23, l01, l23;
                        th::split4<PRE_NONE>(v, 1.0f, h01, h23, l01, l23);
                        const uint32_t off = (uint32_t)t * 64u + ((((uint32_t)k4 >> 1) ^ (((uint32_t)t >> 1) & 3u)) << 4) + ((uint32_t)k4 & 1u) * 8u;
                        *reinterpret_cast<uint2*>(bhi + off) = make_uint2(h01, h23);
                        *reinterpret_cast<uint2*>(bhi + H_TILE + off) = make_uint2(l01, l23);
                    }
                } else {
                float4* bh = reinterpret_cast<float4*>(gen_base + s * STAGE_BYTES + 2 * TILE_BYTES);
                float4* bl = reinterpret_cast<float4*>(gen_base + s * STAGE_BYTES + 3 * TILE_BYTES);
                if (p.gather) {
#pragma unroll
                    for (int i = 0; i < TILE_BYTES / 16 / NUM_XFORM; ++i) {
                        const int item = xt + i * NUM_XFORM;
                        const int t = item >> 3, k4 = item & 7;
                        const int tg = tt * BN + t;
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (tg < p.T) {
                            if (p.gather == 1) {
                                const float* src = span + t * p.hop + kb * BK + k4 * 4;
                                v.x = src[0]; v.y = src[1]; v.z = src[2]; v.w = src[3];
                            } else {   // span too long for the staging buffer (never with the published strides)
                                const float* src = p.wav + (long long)b * p.w_bs + (long long)tg * p.hop + kb * BK + k4 * 4;
                                v.x = __ldg(src); v.y = __ldg(src + 1); v.z = __ldg(src + 2); v.w = __ldg(src + 3);
                            }
                        }
                        float4 l;
                        l.x = tf32_rna(v.x - tf32_trunc(v.x)); l.y = tf32_rna(v.y - tf32_trunc(v.y));
                        l.z = tf32_rna(v.z - tf32_trunc(v.z)); l.w = tf32_rna(v.w - tf32_trunc(v.w));
                        const int dst = t * 8 + (k4 ^ (t & 7));   // float4 index inside the 128B-swizzled K-major tile
                        bh[dst] = v;
                        bl[dst] = l;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < TILE_BYTES / 16 / NUM_XFORM; ++i) {
                        const int idx = xt + i * NUM_XFORM;
                        const float4 v = bh[idx];
                        float4 l;
                        l.x = tf32_rna(v.x - tf32_trunc(v.x)); l.y = tf32_rna(v.y - tf32_trunc(v.y));
                        l.z = tf32_rna(v.z - tf32_trunc(v.z)); l.w = tf32_rna(v.w - tf32_trunc(v.w));
                        bl[idx] = l;
                    }
                }
                }
                fence_proxy_async();
                __syncwarp();
                if ((threadIdx.x & 31) == 0) mbar_arrive(xform_bar(s));   // one arrival per warp (8, not 256, per k-block)
                if (++s == STAGES) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ===================================================================== epilogue: |.|, clamp, log, TMA store
        const int q = warp - 4;
        const int row = q * 32 + lane;            // accumulator row: 2f (re) / 2f+1 (im)
        const int frow = row >> 1;                // bin inside the tile's 64
        // TMA stores / waits of the epilogue: warp q == 0's elected lane (elect.sync picks the same lane every time for the
        // same mask, so the thread that commits a bulk group is the one that waits for it)
        [[maybe_unused]] const bool issuer = (q == 0 && lane == 0);
        const bool even = (lane & 1) == 0;
        const uint32_t sw = p.flat_t ? 0u : (uint32_t)(frow & 7);   // flat tiles stage unswizzled (see gemm_h.cu)
        long long it = 0;
        uint32_t g = 0;
        for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
            const int m_blk = (int)(tile % p.num_m);
            const long long rest = tile / p.num_m;
            const int tt = (int)(rest % p.tiles_t);
            const int b = (int)(rest / p.tiles_t);
            const int acc = (int)(it & 1);
            const uint32_t acc_ph = (uint32_t)((it >> 1) & 1);
            mbar_wait<64>(tfull_bar(acc), acc_ph);
            tc_fence_after();
            const int t0 = tt * BN;
            const int n_chunks = min(BN / 32, (p.T - t0 + 31) / 32);
            const uint32_t t_big = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 2 * BN;
#pragma unroll 1
            for (int c = 0; c < n_chunks; ++c, ++g) {
                const uint32_t obuf = out_base + (g & 1) * OUT_BYTES;
                if (q == 0 && elect_one()) tma_wait_read<1>();
                epi_bar_sync();
                uint32_t rb[32], rs[32];
                tmem_ld32(t_big + c * 32, rb);
                tmem_ld32(t_big + BN + c * 32, rs);
                tmem_ld_wait();
                if (c == n_chunks - 1) {
                    tc_fence_before();
                    mbar_arrive(tempty_bar(acc));
                }
                // Lanes 2p / 2p+1 hold re / im of one bin for all 32 columns.  The pair splits the
                // columns (even lane: 0..15, odd lane: 16..31), swaps what the partner needs with one
                // shuffle per column, and each lane evaluates 16 log-magnitudes:
                //   log(max(sqrt(re^2 + im^2), 1e-5)) = 0.5 * log(max(re^2 + im^2, 1e-10))
                // (no sqrt; ncu showed the sqrt + log of all 128 x 128 accumulators on 4 warps was the
                // bottleneck of the whole kernel).
                float mine[32];
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    mine[j] = kH ? fmaf(__uint_as_float(rs[j]), 1.0f / LO_SCALE, __uint_as_float(rb[j])) * p.c_big
                                 : __uint_as_float(rb[j]) + __uint_as_float(rs[j]);
                const uint32_t orow = obuf + frow * 128;
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                    float o[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int i = j4 * 4 + e;
                        const float send = even ? mine[16 + i] : mine[i];
                        const float recv = __shfl_xor_sync(0xffffffffu, send, 1);
                        const float a = even ? mine[i] : mine[16 + i];   // own component of the column this lane evaluates
                        const float ss = __fadd_rn(__fmul_rn(a, a), __fmul_rn(recv, recv));
                        // 0.5 * ln(ss) = lg2(ss) * (ln 2 / 2).  lg2.approx: absolute error <= 2^-22 on [0.5, 2), <= 2 ulp elsewhere,
                        // i.e. fp32 rounding level like logf -- at 2 issue slots instead of ~25: ncu showed this epilogue
                        // (128 x 128 logarithms per tile on four warps), not the MMAs, bounds the high-rate stages
                        const float sc = fmaxf(ss, 1e-10f);
                        o[e] = p.exact_log ? 0.5f * logf(sc) : __log2f(sc) * 0.34657359027997264f;
                    }
                    const uint32_t chunk = (uint32_t)(even ? j4 : 4 + j4);
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(orow + ((chunk ^ sw) << 4)), "f"(o[0]),
                                 "f"(o[1]), "f"(o[2]), "f"(o[3])
                                 : "memory");
                }
                fence_proxy_async();
                epi_bar_sync();
                if (q == 0 && elect_one()) {
                    if (p.flat_t) tma_store_3d(&map_y, obuf, 0, (t0 + c * 32) / p.flat_t, m_blk * 64);
                    else tma_store_3d(&map_y, obuf, t0 + c * 32, m_blk * 64, b);
                    tma_commit();
                }
            }
        }
        if (q == 0 && elect_one()) tma_wait_all();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        const uint32_t ncols = TMEM_COLS;
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols));
    }
}

}  // namespace stft

// HILCODEC_GEMM=tf32 (the switch of the pointwise GEMMs) also keeps this kernel on 3xTF32
static bool stft_use_h() {
    static const bool v = []() {
        const char* e = std::getenv("HILCODEC_GEMM");
        const char* f = std::getenv("HILCODEC_STFT");   // HILCODEC_STFT=tf32: this kernel only (A/B knob)
        return !(e && std::strcmp(e, "tf32") == 0) && !(f && std::strcmp(f, "tf32") == 0);
    }();
    return v;
}

// windows per clip of a flat launch (0: not flat): short chunks of many streams, as in gemm_h.cu -- T itself, or 4 for
// 1 - 3 windows when the output rows have pitch 4 (the windows past the clip's last one are out of the tensor map's
// bounds and load as zeros)
static inline int stft_flat_t(int T, int y_rs, int B) {
    const int tf = (T < 4 && y_rs == 4) ? 4 : T;
    return (T < 32 && tc_flat_ok(B, tf)) ? tf : 0;
}

bool stft_tc_usable(const PackedMat& Wdft, const float* wav, long long w_bs, int T, const float* Y, long long y_bs,
                    int y_rs, int B, int hop) {
    if (!Wdft.A_hi || !Wdft.A_lo) return false;
    // >= 64 windows per clip, or >= 32 when there are more columns than the skinny FP32 kernel takes (64 streams x 40
    // windows: one partly filled 128-window tile per clip instead of gemm.cu, 43 -> ~12 us)
    if ((Wdft.K % 32) != 0) return false;
    if (T < 64 && !(T >= 32 && (long long)B * T > 512) && !((hop % 4) == 0 && stft_flat_t(T, y_rs, B))) return false;
    if ((w_bs & 3) || (y_rs & 3) || (y_bs & 3)) return false;
    if ((reinterpret_cast<uintptr_t>(wav) & 15) || (reinterpret_cast<uintptr_t>(Y) & 15)) return false;
    return true;
}

cudaError_t launch_stft_tc(const PackedMat& Wdft, const float* wav, long long w_bs, int hop, int B, int T, float* Y,
                           long long y_bs, int y_rs, cudaStream_t st) {
    using namespace stft;
    if (B == 0 || T == 0) return cudaSuccess;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(stft_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(stft_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const bool use_h = stft_use_h() && Wdft.H_hi && Wdft.H_lo;
    const int F = Wdft.M / 2;
    CUtensorMap map_hi, map_lo, map_x, map_y;
    if (use_h) {   // fp16 hi / lo planes of the basis, K-major, 64-byte rows (as gemm_h.cu's A operand)
        const cuuint64_t dims[2] = {(cuuint64_t)Wdft.Kp32, (cuuint64_t)Wdft.Mp128};
        const cuuint64_t strides[1] = {(cuuint64_t)Wdft.Kp32 * 2};
        const cuuint32_t box[2] = {BK, BM};
        if (!make_map_dt(&map_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, Wdft.H_hi, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B) ||
            !make_map_dt(&map_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, Wdft.H_lo, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B))
            return cudaErrorInvalidValue;
    } else {
        const cuuint64_t dims[2] = {(cuuint64_t)Wdft.Kp32, (cuuint64_t)Wdft.Mp128};
        const cuuint64_t strides[1] = {(cuuint64_t)Wdft.Kp32 * 4};
        const cuuint32_t box[2] = {BK, BM};
        if (!make_map(&map_hi, Wdft.A_hi, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B) ||
            !make_map(&map_lo, Wdft.A_lo, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))
            return cudaErrorInvalidValue;
    }
    int gather = (hop % 4) != 0;
    const int flat_t = (T < 32 && !gather) ? stft_flat_t(T, y_rs, B) : 0;
    if (T < 32 && !flat_t) return cudaErrorInvalidValue;
    if (!gather) {
        // im2col as a tensor map with overlapping rows: row t = wav[t*hop .. t*hop + n_fft); a flat box is 128 / flat_t
        // whole clips x their windows: the same [128 windows][32 k] image
        const cuuint64_t dims[3] = {(cuuint64_t)Wdft.K, (cuuint64_t)T, (cuuint64_t)B};
        const cuuint64_t strides[2] = {(cuuint64_t)hop * 4, (cuuint64_t)w_bs * 4};
        const cuuint32_t box[3] = {BK, flat_t ? (cuuint32_t)flat_t : BN, flat_t ? (cuuint32_t)(BN / flat_t) : 1u};
        if (!make_map(&map_x, wav, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) {
            if (flat_t) return cudaErrorInvalidValue;
            gather = 1;
        }
    }
    if (gather) map_x = map_hi;  // unused placeholder
    if (flat_t) {   // {t, clip, f}: one store = 32 / flat_t whole clips x 64 bins, staged unswizzled
        const cuuint64_t dims[3] = {(cuuint64_t)flat_t, (cuuint64_t)B, (cuuint64_t)F};
        const cuuint64_t strides[2] = {(cuuint64_t)y_bs * 4, (cuuint64_t)y_rs * 4};
        const cuuint32_t box[3] = {(cuuint32_t)flat_t, (cuuint32_t)(32 / flat_t), 64};
        if (!make_map(&map_y, Y, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return cudaErrorInvalidValue;
    } else {
        const cuuint64_t dims[3] = {(cuuint64_t)T, (cuuint64_t)F, (cuuint64_t)B};
        const cuuint64_t strides[2] = {(cuuint64_t)y_rs * 4, (cuuint64_t)y_bs * 4};
        const cuuint32_t box[3] = {32, 64, 1};
        if (!make_map(&map_y, Y, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return cudaErrorInvalidValue;
    }
    Params p{};
    p.M = Wdft.M; p.F = F; p.K = Wdft.K; p.T = T; p.B = B; p.hop = hop;
    p.num_m = (Wdft.M + BM - 1) / BM;
    p.tiles_t = (T + BN - 1) / BN;
    p.total_tiles = (long long)p.num_m * p.tiles_t * B;
    if (flat_t) {
        p.flat_t = flat_t; p.T = B * flat_t; p.B = 1;
        p.tiles_t = (p.T + BN - 1) / BN;
        p.total_tiles = (long long)p.num_m * p.tiles_t;
    }
    p.gather = gather ? (((BN - 1) * hop + Wdft.K <= SPAN_FLOATS) ? 1 : 2) : 0; p.wav = wav; p.w_bs = w_bs;
    static const int exact_log = []() { const char* e = std::getenv("HILCODEC_STFT_LOGF"); return (e && e[0] == '1') ? 1 : 0; }();
    p.exact_log = exact_log;
    p.c_big = Wdft.h_inv_scale;
    const int num_sms = device_sm_count();
    const unsigned grid = (unsigned)(p.total_tiles < num_sms ? p.total_tiles : num_sms);
    if (use_h) stft_tc_kernel<true><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(map_hi, map_lo, map_x, map_y, p);
    else stft_tc_kernel<false><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(map_hi, map_lo, map_x, map_y, p);
    return cudaGetLastError();
}

}  // namespace hil
