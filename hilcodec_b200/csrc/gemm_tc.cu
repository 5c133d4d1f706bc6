// Tensor-core (tcgen05 / TMEM / TMA) pointwise GEMM with fp32-level accuracy (3xTF32).
//
//   Y[b][m][t] = sum_k W[m][k] * pre(X[b][k][t]) (+bias[m]) (+R[b][m][t])
//
// Why 3xTF32: VQ indices must match the fp32 reference bit for bit, and a single TF32 pass
// (10-bit mantissa) flips indices.  Every operand is split x = hi + lo with
// hi = tf32(x), lo = tf32(x - hi); D += A_lo*B_hi + A_hi*B_lo + A_hi*B_hi accumulates in
// fp32 in TMEM and drops only the ~2^-22 lo*lo term.  For the activations hi = trunc(x) is what
// the tensor core reads out of a raw fp32 container, so the raw tile doubles as B_hi.
// Per k-step (K = 8) two MMAs are issued: A_hi x [B_hi | B_lo] as one N = 256 instruction into
// the adjacent [big | small] accumulators, then A_lo x B_hi (N = 128) into `small` -- the kernel
// is shared-memory-bandwidth bound (SS-mode operands), so reading A_hi once instead of twice
// is worth 10 %.
//
// Mapping (one persistent CTA per SM, 128 x 128 output tile, BK = 32):
//   A = weights  [M = Cout rows, K]   K-major, SWIZZLE_128B, hi/lo pre-split at finalize, TMA 2D
//   B = activations [K rows, N = time] MN-major (time contiguous, NCW), SWIZZLE_128B_ATOM_32B, TMA 3D
//       boxes of 32 k x 32 t; raw fp32 lands in the B_hi slot, 8 transform warps apply the
//       ELU prologue in place and write B_lo next to it (elementwise, so the swizzle is untouched)
//   D in TMEM: per tile two 128-column accumulators -- "big" (A_hi*B_hi) and "small" (the two
//       cross terms) -- so the big sum is rounded once per k-step instead of three times (the
//       tensor core truncates after every accumulate); both are double buffered (4 x 128 = all
//       512 columns), so the epilogue of tile i overlaps the MMAs of tile i+1.  Lane = Cout
//       row, column = time: bias is a per-thread scalar.
//   Epilogue: TMEM -> registers -> (big + small + bias) -> 128B-swizzled smem staging ->
//       TMA tensor store (or TMA reduce-add when the 1x1 output is added in place to h, the
//       SpecBlock case); out-of-range rows / columns are clipped by the tensor map.
// Warp roles: 0 TMA producer | 1 MMA issuer | 2 TMEM alloc | 4-7 epilogue | 8-15 transform.
#include "tc_ptx.cuh"

namespace hil {
namespace tc {

constexpr int BM = 128, BN = 128, BK = 32;
constexpr int STAGES = 3;
constexpr int TILE_BYTES = BM * BK * 4;         // 16 KB: one operand tile
constexpr int PANEL_BYTES = BK * 32 * 4;        // 4 KB: 32 k-rows x 32 t (one TMA box of B)
constexpr int STAGE_BYTES = 4 * TILE_BYTES;     // A_hi, A_lo, B_hi, B_lo
constexpr int STAGE_TX = 3 * TILE_BYTES;        // bytes TMA writes per stage
constexpr int NUM_THREADS = 512;
constexpr int NUM_XFORM = 256;
constexpr int NUM_EPI = 128;
constexpr int OUT_BYTES = BM * 32 * 4;           // 16 KB: one 128-row x 32-column output chunk
constexpr int TMEM_COLS = 512;
constexpr size_t SMEM_BYTES = 1024 + (size_t)STAGES * STAGE_BYTES + 2 * OUT_BYTES + 256;

struct Params {
    int M, K, T, B;
    int num_m, tiles_t;
    long long total_tiles;
    int pre;
    float pre_scale;
    const float* bias;
    int reduce_add;  // 1: Y += tile (TMA reduce), 0: Y = tile
    // time tiling: tile tt covers columns [tt * t_step - t_halo, ... + BN)
    int t_step, t_halo;
    // fused DWS epilogue (kDw): y (+)= bias_dw + dw5(pw), causal with cache; (+)= when reduce_add
    const float* dw_w;       // [M][5]
    const float* dw_b;       // [M] or null
    const float* cache_in;   // [B][M][4]
    float* cache_out;        // [B][M][4]
};

constexpr uint32_t IDESC = make_idesc(BM, BN, 1);  // kind::tf32, D = f32, A K-major, B MN-major
constexpr uint32_t IDESC_N256 = make_idesc(BM, 2 * BN, 1);

// ------------------------------------------------------------------------------- kernel
template <bool kDw>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
               const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_y,
               const __grid_constant__ CUtensorMap map_y28, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t out_base = base + STAGES * STAGE_BYTES;
    const uint32_t bars = out_base + 2 * OUT_BYTES;
    // barrier slots (8 bytes each)
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto xform_bar = [&](int s) { return bars + 8u * (STAGES + s); };
    auto empty_bar = [&](int s) { return bars + 8u * (2 * STAGES + s); };
    auto tfull_bar = [&](int a) { return bars + 8u * (3 * STAGES + a); };
    auto tempty_bar = [&](int a) { return bars + 8u * (3 * STAGES + 2 + a); };
    const uint32_t tmem_slot = bars + 8u * (3 * STAGES + 4);
    constexpr int kNumXform = NUM_XFORM;
    uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb = (p.K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_a_hi);
        prefetch_tmap(&map_a_lo);
        prefetch_tmap(&map_x);
        prefetch_tmap(&map_y);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(xform_bar(s), kNumXform);
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
        if (lane == 0) {
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
                    mbar_arrive_expect_tx(full_bar(s), STAGE_TX);
                    tma_load_2d(&map_a_hi, st, full_bar(s), kb * BK, m_blk * BM);
                    tma_load_2d(&map_a_lo, st + TILE_BYTES, full_bar(s), kb * BK, m_blk * BM);
#pragma unroll
                    for (int pnl = 0; pnl < 4; ++pnl)
                        tma_load_3d(&map_x, st + 2 * TILE_BYTES + pnl * PANEL_BYTES, full_bar(s),
                                    tt * p.t_step - p.t_halo + pnl * 32, kb * BK, b);
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
                mbar_wait(full_bar(s), ph);
                mbar_wait(xform_bar(s), ph);
                tc_fence_after();
                // The tensor core reads only the top 19 bits of a tf32 operand, so the raw fp32 tile IS the
                // hi operand (hi = trunc(x)).  B_hi and B_lo sit back to back in the stage (8 panels of 32
                // columns), so ONE N = 256 MMA computes A_hi*[B_hi | B_lo] into the adjacent [big | small]
                // accumulators and a second N = 128 MMA adds A_lo*B_hi to `small`: A_hi is read from shared
                // memory once per k-step instead of twice (shared-memory bandwidth is what bounds this kernel).
                if (lane == 0) {
#pragma unroll
                    for (int j = 0; j < BK / 8; ++j) {
                        // A: 8-row groups 1024 B apart; +32 B walks K inside the 128-byte swizzle row.
                        // B: 32-column panels PANEL_BYTES apart (LBO), 4-row swizzle groups 512 B apart
                        //    (SBO); one MMA (K = 8) consumes 8 k-rows = 1024 B.
                        const uint64_t a_hi = make_desc(st + j * 32, 16, 1024, 2);
                        const uint64_t a_lo = make_desc(st + TILE_BYTES + j * 32, 16, 1024, 2);
                        const uint64_t b_hl = make_desc(st + 2 * TILE_BYTES + j * 1024, PANEL_BYTES, 512, 1);
                        umma_tf32(d_big, a_hi, b_hl, IDESC_N256, (kb | j) != 0);
                        umma_tf32(d_small, a_lo, b_hl, IDESC, 1);
                    }
                    umma_commit(empty_bar(s));
                    if (kb == nkb - 1) umma_commit(tfull_bar(acc));
                }
                __syncwarp();
                if (++s == STAGES) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp >= 8 && warp < 8 + kNumXform / 32) {
        // ===================================================================== transform (ELU + hi/lo split)
        const int xt = threadIdx.x - 256;
        int s = 0;
        uint32_t ph = 0;
        for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(full_bar(s), ph);
                float4* bh = reinterpret_cast<float4*>(gen_base + s * STAGE_BYTES + 2 * TILE_BYTES);
                float4* bl = reinterpret_cast<float4*>(gen_base + s * STAGE_BYTES + 3 * TILE_BYTES);
                if (p.pre == PRE_NONE) {
#pragma unroll
                    for (int i = 0; i < TILE_BYTES / 16 / kNumXform; ++i) {
                        const int idx = xt + i * kNumXform;
                        const float4 v = bh[idx];
                        float4 l;
                        l.x = tf32_rna(v.x - tf32_trunc(v.x)); l.y = tf32_rna(v.y - tf32_trunc(v.y));
                        l.z = tf32_rna(v.z - tf32_trunc(v.z)); l.w = tf32_rna(v.w - tf32_trunc(v.w));
                        bl[idx] = l;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < TILE_BYTES / 16 / kNumXform; ++i) {
                        const int idx = xt + i * kNumXform;
                        float4 v = bh[idx];  // pre_scale is 1.0 for PRE_ELU (x * 1.0f is exact)
                        v.x = elu_fast(v.x * p.pre_scale); v.y = elu_fast(v.y * p.pre_scale);
                        v.z = elu_fast(v.z * p.pre_scale); v.w = elu_fast(v.w * p.pre_scale);
                        float4 l;
                        l.x = tf32_rna(v.x - tf32_trunc(v.x)); l.y = tf32_rna(v.y - tf32_trunc(v.y));
                        l.z = tf32_rna(v.z - tf32_trunc(v.z)); l.w = tf32_rna(v.w - tf32_trunc(v.w));
                        bh[idx] = v;
                        bl[idx] = l;
                    }
                }
                fence_proxy_async();
                mbar_arrive(xform_bar(s));
                if (++s == STAGES) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ===================================================================== epilogue
        if constexpr (!kDw) {
            const int q = warp - 4;
            const int row = q * 32 + lane;                      // row inside the 128-row tile = TMEM lane
            const bool issuer = (q == 0 && lane == 0);
            const uint32_t sw = (uint32_t)(row & 7);            // 128B-swizzle phase of this row
            long long it = 0;
            uint32_t g = 0;                                      // running chunk counter -> staging buffer parity
            for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
                const int m_blk = (int)(tile % p.num_m);
                const long long rest = tile / p.num_m;
                const int tt = (int)(rest % p.tiles_t);
                const int b = (int)(rest / p.tiles_t);
                const int acc = (int)(it & 1);
                const uint32_t acc_ph = (uint32_t)((it >> 1) & 1);
                mbar_wait<64>(tfull_bar(acc), acc_ph);
                tc_fence_after();
                const int m = m_blk * BM + row;
                const float bv = (m < p.M && p.bias) ? p.bias[m] : 0.f;
                const int t0 = tt * BN;
                const int n_chunks = min(BN / 32, (p.T - t0 + 31) / 32);
                const uint32_t t_big = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 2 * BN;
    #pragma unroll 1
                for (int c = 0; c < n_chunks; ++c, ++g) {
                    const uint32_t obuf = out_base + (g & 1) * OUT_BYTES;
                    if (issuer) tma_wait_read<1>();              // the store that used this buffer two chunks ago has drained it
                    epi_bar_sync();
                    uint32_t rb[32], rs[32];
                    tmem_ld32(t_big + c * 32, rb);
                    tmem_ld32(t_big + BN + c * 32, rs);
                    tmem_ld_wait();
                    if (c == n_chunks - 1) {
                        tc_fence_before();
                        mbar_arrive(tempty_bar(acc));
                    }
                    const uint32_t orow = obuf + row * 128;
    #pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float o0 = (__uint_as_float(rb[4 * j + 0]) + __uint_as_float(rs[4 * j + 0])) + bv;
                        const float o1 = (__uint_as_float(rb[4 * j + 1]) + __uint_as_float(rs[4 * j + 1])) + bv;
                        const float o2 = (__uint_as_float(rb[4 * j + 2]) + __uint_as_float(rs[4 * j + 2])) + bv;
                        const float o3 = (__uint_as_float(rb[4 * j + 3]) + __uint_as_float(rs[4 * j + 3])) + bv;
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(orow + (((uint32_t)j ^ sw) << 4)), "f"(o0),
                                     "f"(o1), "f"(o2), "f"(o3)
                                     : "memory");
                    }
                    fence_proxy_async();
                    epi_bar_sync();
                    if (issuer) {
                        if (p.reduce_add) tma_reduce_add_3d(&map_y, obuf, t0 + c * 32, m_blk * BM, b);
                        else tma_store_3d(&map_y, obuf, t0 + c * 32, m_blk * BM, b);
                        tma_commit();
                    }
                }
            }
            if (issuer) tma_wait_all();

        } else {
            // ---- fused DWS epilogue: pointwise tile -> causal depthwise k5 + bias -> TMA store / reduce-add.
            // The tile holds 128 pointwise columns for times [t0-4, t0+124); each thread owns one
            // channel row and slides the 5-tap window along it in registers (the 4 halo columns come
            // from this tile, or from cache_in at the start of the chunk).  Chunk 0 yields 28 outputs
            // (dense 112-byte rows, box-28 tensor map), chunks 1-3 yield 32 (128B-swizzled rows).
            // With a residual skip the output is accumulated in place (h += ...) by TMA reduce-add.
            const int q = warp - 4;
            const int row = q * 32 + lane;
            const bool issuer = (q == 0 && lane == 0);
            const uint32_t sw = (uint32_t)(row & 7);
            long long it = 0;
            uint32_t g = 0;
            for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
                const int m_blk = (int)(tile % p.num_m);
                const long long rest = tile / p.num_m;
                const int tt = (int)(rest % p.tiles_t);
                const int b = (int)(rest / p.tiles_t);
                const int acc = (int)(it & 1);
                const uint32_t acc_ph = (uint32_t)((it >> 1) & 1);
                const int m = m_blk * BM + row;
                const bool row_ok = m < p.M;
                float wk[5];
#pragma unroll
                for (int k = 0; k < 5; ++k) wk[k] = row_ok ? p.dw_w[m * 5 + k] : 0.f;
                const float bv = (row_ok && p.dw_b) ? p.dw_b[m] : 0.f;
                const int tcol0 = tt * p.t_step - p.t_halo;       // time of tile column 0
                const int n_chunks = min(BN / 32, (p.T - tcol0 + 31) / 32);
                const bool has_tail = row_ok && (tcol0 + BN > p.T - 4);  // tile holds some of the last 4 columns
                float carry[4] = {0.f, 0.f, 0.f, 0.f};
                if (tt == 0 && row_ok) {
                    const float4 cv = *reinterpret_cast<const float4*>(p.cache_in + ((size_t)b * p.M + m) * 4);
                    carry[0] = cv.x; carry[1] = cv.y; carry[2] = cv.z; carry[3] = cv.w;
                }
                mbar_wait<64>(tfull_bar(acc), acc_ph);
                tc_fence_after();
                const uint32_t t_big = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 2 * BN;
#pragma unroll 1
                for (int c = 0; c < n_chunks; ++c, ++g) {
                    const uint32_t obuf = out_base + (g & 1) * OUT_BYTES;
                    uint32_t rb[32], rs[32];
                    tmem_ld32(t_big + c * 32, rb);
                    tmem_ld32(t_big + BN + c * 32, rs);
                    tmem_ld_wait();
                    if (c == n_chunks - 1) {
                        tc_fence_before();
                        mbar_arrive(tempty_bar(acc));
                    }
                    float v[36];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[4 + j] = __uint_as_float(rb[j]) + __uint_as_float(rs[j]);
                    // columns 0..3 of the first tile are times -4..-1: the cache handed in by the caller
                    if (c == 0 && tt == 0) { v[4] = carry[0]; v[5] = carry[1]; v[6] = carry[2]; v[7] = carry[3]; }
                    v[0] = carry[0]; v[1] = carry[1]; v[2] = carry[2]; v[3] = carry[3];
                    if (has_tail) {  // new cache = pointwise outputs at times T-4..T-1 (each time is owned by one tile)
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int col = c * 32 + j;
                            const int t = tcol0 + col;
                            if (col >= p.t_halo && t >= p.T - 4 && t < p.T)
                                p.cache_out[((size_t)b * p.M + m) * 4 + (t - (p.T - 4))] = v[4 + j];
                        }
                    }
                    if (issuer) tma_wait_read<1>();   // the store that used this buffer two chunks ago has drained it
                    epi_bar_sync();
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {
                        float o[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int i = j4 * 4 + e;  // output i of this chunk uses v[i..i+4]
                            float a = 0.f;
#pragma unroll
                            for (int k = 0; k < 5; ++k) a = fmaf(wk[k], v[i + k], a);
                            o[e] = a + bv;
                        }
                        uint32_t dst;
                        if (c == 0) {
                            if (j4 == 0) continue;   // outputs 0..3 of chunk 0 belong to the previous tile
                            dst = obuf + row * 112 + (j4 - 1) * 16;
                        } else {
                            dst = obuf + row * 128 + (((uint32_t)j4 ^ sw) << 4);
                        }
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "f"(o[0]), "f"(o[1]), "f"(o[2]),
                                     "f"(o[3])
                                     : "memory");
                    }
                    carry[0] = v[32]; carry[1] = v[33]; carry[2] = v[34]; carry[3] = v[35];
                    fence_proxy_async();
                    epi_bar_sync();
                    if (issuer) {
                        const CUtensorMap* mp = c == 0 ? &map_y28 : &map_y;
                        const int tc = c == 0 ? tcol0 + p.t_halo : tcol0 + c * 32;
                        if (p.reduce_add) tma_reduce_add_3d(mp, obuf, tc, m_blk * BM, b);
                        else tma_store_3d(mp, obuf, tc, m_blk * BM, b);
                        tma_commit();
                    }
                }
            }
            if (issuer) tma_wait_all();
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

// ------------------------------------------------------------------------------- host side
}  // namespace tc

bool gemm_tc_usable(const PackedMat& W, const float* X, long long x_bs, int x_rs, int T, const float* R, const float* Y,
                    long long y_bs, int y_rs, int B) {
    if (!W.A_hi || !W.A_lo) return false;
    if (!tc_chunk_ok(B, T)) return false;  // short chunks (streaming) go to the flattened-column FP32 kernels
    if ((x_rs & 3) || (x_bs & 3) || (y_rs & 3) || (y_bs & 3)) return false;
    if ((reinterpret_cast<uintptr_t>(X) & 15) || (reinterpret_cast<uintptr_t>(Y) & 15)) return false;
    if (R && (reinterpret_cast<uintptr_t>(R) & 15)) return false;
    return true;
}

static cudaError_t tc_common(const PackedMat& W, const float* X, long long x_bs, int x_rs, int B, int T, CUtensorMap* map_hi,
                             CUtensorMap* map_lo, CUtensorMap* map_x, int* num_sms_out) {
    using namespace tc;
    static int num_sms = 0;
    static bool attr_set = false;
    if (!num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(gemm_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    *num_sms_out = num_sms;
    {
        const cuuint64_t dims[2] = {(cuuint64_t)W.Kp32, (cuuint64_t)W.Mp128};
        const cuuint64_t strides[1] = {(cuuint64_t)W.Kp32 * 4};
        const cuuint32_t box[2] = {BK, BM};
        if (!make_map(map_hi, W.A_hi, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B) ||
            !make_map(map_lo, W.A_lo, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))
            return cudaErrorInvalidValue;
    }
    {
        const cuuint64_t dims[3] = {(cuuint64_t)T, (cuuint64_t)W.K, (cuuint64_t)B};
        const cuuint64_t strides[2] = {(cuuint64_t)x_rs * 4, (cuuint64_t)x_bs * 4};
        const cuuint32_t box[3] = {32, BK, 1};
        if (!make_map(map_x, X, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return cudaErrorInvalidValue;
    }
    return cudaSuccess;
}

cudaError_t launch_gemm_tc(const PackedMat& W, const float* X, long long x_bs, int x_rs, int B, int T, int pre,
                           float pre_scale, const float* bias, const float* R, float* Y, long long y_bs, int y_rs,
                           cudaStream_t st) {
    using namespace tc;
    if (B == 0 || T == 0) return cudaSuccess;
    CUtensorMap map_hi, map_lo, map_x, map_y;
    int num_sms = 0;
    cudaError_t e = tc_common(W, X, x_bs, x_rs, B, T, &map_hi, &map_lo, &map_x, &num_sms);
    if (e != cudaSuccess) return e;
    {
        const cuuint64_t dims[3] = {(cuuint64_t)T, (cuuint64_t)W.M, (cuuint64_t)B};
        const cuuint64_t strides[2] = {(cuuint64_t)y_rs * 4, (cuuint64_t)y_bs * 4};
        const cuuint32_t box[3] = {32, BM, 1};
        if (!make_map(&map_y, Y, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return cudaErrorInvalidValue;
    }
    if (R && R != Y) {  // out-of-place residual: seed Y with R, then accumulate in place
        for (int b = 0; b < B; ++b) {
            e = cudaMemcpy2DAsync(Y + (long long)b * y_bs, (size_t)y_rs * 4, R + (long long)b * y_bs, (size_t)y_rs * 4,
                                  (size_t)T * 4, W.M, cudaMemcpyDeviceToDevice, st);
            if (e != cudaSuccess) return e;
        }
    }
    Params p{};
    p.M = W.M; p.K = W.K; p.T = T; p.B = B;
    p.num_m = (W.M + BM - 1) / BM;
    p.t_step = BN; p.t_halo = 0;
    p.tiles_t = (T + p.t_step - 1) / p.t_step;
    p.total_tiles = (long long)p.num_m * p.tiles_t * B;
    p.pre = pre; p.pre_scale = (pre == PRE_SCALE_ELU) ? pre_scale : 1.0f; p.bias = bias; p.reduce_add = R ? 1 : 0;
    const unsigned grid = (unsigned)(p.total_tiles < num_sms ? p.total_tiles : num_sms);
    gemm_tc_kernel<false><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(map_hi, map_lo, map_x, map_y, map_y, p);
    return cudaGetLastError();
}

// Fused DWSBlock (streaming.py:189-192): y = post(dw5(W * pre(x)) + b_dw (+ skip)), with the causal
// cache of the depthwise conv (the last 4 pointwise outputs) read from cache_in / written to cache_out.
cudaError_t launch_gemm_tc_dw(const PackedMat& W, const float* X, long long x_bs, int x_rs, int B, int T, int pre,
                              float pre_scale, const float* dw_w, const float* dw_b, const float* cache_in,
                              float* cache_out, const float* skip, float* Y, long long y_bs, int y_rs, cudaStream_t st) {
    using namespace tc;
    if (B == 0 || T == 0) return cudaSuccess;
    CUtensorMap map_hi, map_lo, map_x, map_y, map_y28;
    int num_sms = 0;
    cudaError_t e = tc_common(W, X, x_bs, x_rs, B, T, &map_hi, &map_lo, &map_x, &num_sms);
    if (e != cudaSuccess) return e;
    {
        const cuuint64_t dims[3] = {(cuuint64_t)T, (cuuint64_t)W.M, (cuuint64_t)B};
        const cuuint64_t strides[2] = {(cuuint64_t)y_rs * 4, (cuuint64_t)y_bs * 4};
        const cuuint32_t box[3] = {32, BM, 1};
        const cuuint32_t box28[3] = {28, BM, 1};
        if (!make_map(&map_y, Y, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B) ||
            !make_map(&map_y28, Y, 3, dims, strides, box28, CU_TENSOR_MAP_SWIZZLE_NONE))
            return cudaErrorInvalidValue;
    }
    if (skip && skip != Y) {  // out-of-place residual: seed Y with the skip, then accumulate in place
        for (int b = 0; b < B; ++b) {
            e = cudaMemcpy2DAsync(Y + (long long)b * y_bs, (size_t)y_rs * 4, skip + (long long)b * y_bs, (size_t)y_rs * 4,
                                  (size_t)T * 4, W.M, cudaMemcpyDeviceToDevice, st);
            if (e != cudaSuccess) return e;
        }
    }
    Params p{};
    p.M = W.M; p.K = W.K; p.T = T; p.B = B;
    p.num_m = (W.M + BM - 1) / BM;
    p.t_step = BN - 4; p.t_halo = 4;
    p.tiles_t = (T + p.t_step - 1) / p.t_step;
    p.total_tiles = (long long)p.num_m * p.tiles_t * B;
    p.pre = pre; p.pre_scale = (pre == PRE_SCALE_ELU) ? pre_scale : 1.0f;
    p.dw_w = dw_w; p.dw_b = dw_b; p.cache_in = cache_in; p.cache_out = cache_out;
    p.reduce_add = skip ? 1 : 0;
    const unsigned grid = (unsigned)(p.total_tiles < num_sms ? p.total_tiles : num_sms);
    gemm_tc_kernel<true><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(map_hi, map_lo, map_x, map_y, map_y28, p);
    return cudaGetLastError();
}

}  // namespace hil
