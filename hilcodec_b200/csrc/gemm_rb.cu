// Fused ResBlock on the tensor pipe: both DWSBlocks and the residual add in ONE kernel.
//
//   ResBlock.forward (streaming.py:252-275, residual scale folded into the second depthwise):
//     h <- h + dw5_1(W1 * ELU(dw5_0(W0 * ELU(pre * h)) + b0)) + b1
//
// With two fused-DWS launches (gemm_h.cu) a ResBlock moves 5 tensors through HBM (read h, write a, read a,
// read-modify-write h); for the high-rate stages (C <= 256, T >= 3000) that traffic is what bounds them.  Here
// a CTA keeps the whole channel column of a time tile on chip, so HBM sees one read and one write of h:
//
//   G1   D1[C][BN] = W0 * split(ELU(pre * h[:, tile]))            tcgen05 kind::f16, 3 MMAs per product (gemm_h.cu)
//   E1   a = dw5_0(D1) + b0;  B2 = split(ELU(a))                  TMEM -> registers -> fp16 hi/lo operand in SMEM
//   G2   D2[C][BN] = W1 * B2                                      B operand never leaves the SM
//   E2   y = dw5_1(D2) + b1;  h[:, tile] += y                     TMA reduce-add, in place
//
// The tile carries 8 halo columns (two causal k5 depthwise convs), so BN columns produce BN - 8 outputs.
// In-place safety: a tile's halo columns belong to its left neighbour, which may already have been updated,
// so a small gather kernel copies the 8 columns in front of every tile to a side buffer first (6.7 % / 14 %
// of one read) and the tile loads [halo | own columns] with two TMA boxes; its own columns are read before it
// writes them and nobody else touches them.
//
// Geometry: C <= 128: one 128-row m-block, BN = 128: an accumulator (big | small) fills 256 TMEM columns, D1 and D2
// fill the 512.  (The kernel stays templated on BN; the two-m-block BN = 64 instantiation for 128 < C <= 256
// re-streams both weight matrices from L2 for every 56 outputs and issues half-width MMAs: measured slower than two
// fused-DWS launches in round 1 and again at the end of round 2 (+2.3 ms per step), so it is opt-in: HILCODEC_RB_WIDE=1.)
// Warp roles (512 threads, 1 CTA/SM): 0 X producer (TMA), 1 MMA issuer, 2 TMEM allocator, 3 weight producer
// (TMA, W0 then W1 pieces through one ring), 8-15 workers: transform (ELU + split of the input tile) and E1, two warps
// per TMEM lane quarter; 4-7 epilogue E2.  E2 of tile i overlaps the transform / G1 / E1 of tile i+1.
// (Tried and measured slower: issuing G1(i+1) ahead of G2(i); per-warp 32-row staging + TMA stores in E2.)
// Arithmetic is the one of gemm_h.cu, instruction for instruction, so the result is bit-identical to the two
// fused-DWS launches it replaces (tests/test_gpu_ops.py::test_resblock_fused).
#include <cstdlib>

#include "h_split.cuh"

namespace hil {
namespace rb {

using namespace th;

constexpr int BM = 128, BK = 32, HALO = 8;
constexpr int NUM_THREADS = 512;
constexpr int NUM_XFORM_WARPS = 8;
constexpr int NUM_EPI = 128;
constexpr int XG = 2;                             // k-blocks the worker warps convert concurrently
constexpr int XW_PER_G = NUM_XFORM_WARPS / XG;
// Stage counts are multiples of XG so that a ring stage is always filled for, and released by, the same transform
// group: the groups wait by mbarrier parity, which is ambiguous if a group could run two uses of a stage ahead of the
// other group (e.g. 3 raw stages shared by 2 groups: a late TMA completion for k-block n-3 would let the other group
// pass the wait for k-block n).  4 raw + 2 weight stages cost the same shared memory as 3 + 3.
constexpr int RAW_STAGES = 4, B_STAGES = 2, A_STAGES = 2;
static_assert(RAW_STAGES % XG == 0 && B_STAGES % XG == 0, "see above");
constexpr int A_TILE = BM * BK * 2;       // 8 KB: one fp16 plane of a 128 x 32 weight piece
constexpr int A_PIECE = 2 * A_TILE;       // A_hi | A_lo
constexpr int B_PANEL = 8 * 128 * (BK / 8);  // 4 KB: 64 columns x 32 k (4 swizzle atoms of 8 k-rows x 128 B)
constexpr int OUT_BYTES = BM * 32 * 4;    // 16 KB staging chunk: 128 rows x 32 columns fp32
constexpr int B2_BYTES = 64 * 1024;       // the second GEMM's whole B operand: C x BN x (hi + lo) fp16, C * BN <= 16384
constexpr int TMEM_COLS = 512;
#ifndef RB_HOIST_TAPS
#define RB_HOIST_TAPS 1
#endif
constexpr bool RB_HOIST = RB_HOIST_TAPS != 0;

template <int BN>
struct Geo {
    static constexpr int NUM_M = 128 / BN;            // m-blocks per tile
    static constexpr int VAL = BN - HALO;             // output columns per tile
    static constexpr int RAW_MAIN = BK * VAL * 4;     // own columns, dense [32][VAL] fp32
    static constexpr int RAW_BYTES = BK * BN * 4;     // + [32][8] halo columns behind it
    static constexpr int B_TILE = BK * BN * 2;        // one fp16 plane of a k-block of B
    static constexpr int B_STAGE = 2 * B_TILE;        // B_hi panels | B_lo panels
    static constexpr size_t SMEM = 1024 + (size_t)RAW_STAGES * RAW_BYTES + (size_t)B_STAGES * B_STAGE +
                                   (size_t)A_STAGES * A_PIECE + B2_BYTES + 2 * OUT_BYTES + 256;
    static constexpr uint32_t IDESC_2N = make_idesc_f16(BM, 2 * BN);
    static constexpr uint32_t IDESC_N = make_idesc_f16(BM, BN);
};

struct Params {
    int C, T, B;
    int nkb, num_m, tiles_t;
    long long total_tiles;
    int pre;
    int xform_sleep;               // ns of back-off in the worker warps' barrier polls (0 = spin)
    float pre_scale;
    float c_big0, c_big1;          // 2^-s of W0 / W1 (fp16 weight scaling, see gemm_h.cu)
    const float* dw0_w;            // [C][5]
    const float* dw0_b;            // [C] or null
    const float* dw1_w;
    const float* dw1_b;
    const float* c0_in;            // [B][C][4]: last 4 pointwise-0 outputs of the previous chunk
    float* c0_out;
    const float* c1_in;            // same for pointwise-1
    float* c1_out;
};

// 8 depthwise outputs -> ELU -> fp16 hi/lo words (4 + 4)
__device__ __forceinline__ void elu_split8(const float (&o)[8], uint32_t (&hi)[4], uint32_t (&lo)[4]) {
    split4<PRE_ELU>(make_float4(o[0], o[1], o[2], o[3]), 1.0f, hi[0], hi[1], lo[0], lo[1]);
    split4<PRE_ELU>(make_float4(o[4], o[5], o[6], o[7]), 1.0f, hi[2], hi[3], lo[2], lo[3]);
}

template <int kPre, int BN>
__device__ __forceinline__ void xform_tile_rows(const uint8_t* raw, uint32_t bdst, int xw, int lane, float s) {
    using G = Geo<BN>;
    constexpr int LPR = BN / 4;       // lanes per k-row (a lane owns 4 consecutive columns)
    constexpr int RPI = 32 / LPR;     // k-rows per warp iteration
    constexpr int NIT = 4 / RPI;      // a warp owns 4 k-rows of the box
    const int cg = lane % LPR;
    float4 v[NIT];
#pragma unroll
    for (int q = 0; q < NIT; ++q) {
        const int k = xw * 4 + q * RPI + lane / LPR;
        // tile columns 0..7 are the halo box (behind the main box), 8.. the tile's own columns
        const uint8_t* src = cg < 2 ? raw + G::RAW_MAIN + k * (HALO * 4) + cg * 16 : raw + k * (G::VAL * 4) + (cg - 2) * 16;
        v[q] = *reinterpret_cast<const float4*>(src);
    }
    const uint32_t panel = (uint32_t)(cg >> 4), chunk = (uint32_t)((cg & 15) >> 1), half8 = (uint32_t)(cg & 1) * 8u;
#pragma unroll
    for (int q = 0; q < NIT; ++q) {
        const uint32_t k = (uint32_t)(xw * 4 + q * RPI + lane / LPR);
        uint32_t h01, h23, l01, l23;
        split4<kPre>(v[q], s, h01, h23, l01, l23);
        const uint32_t dst = bdst + panel * B_PANEL + (k >> 3) * 1024u + (k & 7u) * 128u + ((chunk ^ (k & 7u)) << 4) + half8;
        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(dst), "r"(h01), "r"(h23) : "memory");
        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(dst + (uint32_t)G::B_TILE), "r"(l01), "r"(l23) : "memory");
    }
}

// ------------------------------------------------------------------------------- kernel
template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
resblock_kernel(const __grid_constant__ CUtensorMap map_a0_hi, const __grid_constant__ CUtensorMap map_a0_lo,
                const __grid_constant__ CUtensorMap map_a1_hi, const __grid_constant__ CUtensorMap map_a1_lo,
                const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_halo,
                const __grid_constant__ CUtensorMap map_y, const __grid_constant__ CUtensorMap map_y24, const Params p) {
    using G = Geo<BN>;
    constexpr int NUM_M = G::NUM_M;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t raw_base = base;
    const uint32_t b_base = raw_base + RAW_STAGES * G::RAW_BYTES;
    const uint32_t a_base = b_base + B_STAGES * G::B_STAGE;
    const uint32_t b2_base = a_base + A_STAGES * A_PIECE;
    const uint32_t out_base = b2_base + B2_BYTES;
    const uint32_t bars = out_base + 2 * OUT_BYTES;
    auto raw_full = [&](int r) { return bars + 8u * r; };
    auto raw_empty = [&](int r) { return bars + 8u * (RAW_STAGES + r); };
    auto b_ready = [&](int s) { return bars + 8u * (2 * RAW_STAGES + s); };
    auto b_empty = [&](int s) { return bars + 8u * (2 * RAW_STAGES + B_STAGES + s); };
    auto a_full = [&](int s) { return bars + 8u * (2 * RAW_STAGES + 2 * B_STAGES + s); };
    auto a_empty = [&](int s) { return bars + 8u * (2 * RAW_STAGES + 2 * B_STAGES + A_STAGES + s); };
    constexpr int NB = 2 * RAW_STAGES + 2 * B_STAGES + 2 * A_STAGES;
    const uint32_t d1_full = bars + 8u * NB, d1_empty = bars + 8u * (NB + 1), b2_ready = bars + 8u * (NB + 2),
                   d2_full = bars + 8u * (NB + 3), d2_empty = bars + 8u * (NB + 4);
    const uint32_t tmem_slot = bars + 8u * (NB + 5);
    uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb = p.nkb;
    const uint32_t n_my = (uint32_t)((p.total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);   // tiles of this CTA

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_a0_hi); prefetch_tmap(&map_a0_lo); prefetch_tmap(&map_a1_hi); prefetch_tmap(&map_a1_lo);
        prefetch_tmap(&map_x); prefetch_tmap(&map_halo); prefetch_tmap(&map_y); prefetch_tmap(&map_y24);
    }
    if (warp == 1 && lane == 0) {
        for (int r = 0; r < RAW_STAGES; ++r) {
            mbar_init(raw_full(r), 1);
            mbar_init(raw_empty(r), XW_PER_G);
        }
        for (int s = 0; s < B_STAGES; ++s) {
            mbar_init(b_ready(s), XW_PER_G);
            mbar_init(b_empty(s), 1);
        }
        for (int s = 0; s < A_STAGES; ++s) {
            mbar_init(a_full(s), 1);
            mbar_init(a_empty(s), 1);
        }
        mbar_init(d1_full, 1);
        mbar_init(d1_empty, NUM_XFORM_WARPS * 32);
        mbar_init(b2_ready, NUM_XFORM_WARPS * 32);
        mbar_init(d2_full, 1);
        mbar_init(d2_empty, NUM_EPI);
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
        // ===================================================================== X producer: [own columns | halo] boxes
        if (elect_one()) {
            int r = 0;
            uint32_t ph = 0;
            for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const int tt = (int)(tile % p.tiles_t);
                const int b = (int)(tile / p.tiles_t);
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait<32>(raw_empty(r), ph ^ 1);
                    mbar_arrive_expect_tx(raw_full(r), G::RAW_BYTES);
                    const uint32_t dst = raw_base + r * G::RAW_BYTES;
                    tma_load_3d(&map_x, dst, raw_full(r), tt * G::VAL, kb * BK, b);
                    tma_load_3d(&map_halo, dst + G::RAW_MAIN, raw_full(r), tt * HALO, kb * BK, b);
                    if (++r == RAW_STAGES) { r = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 3) {
        // ===================================================================== weight producer (same order as the MMA issuer)
        if (elect_one()) {
            int s = 0;
            uint32_t ph = 0;
            auto pieces = [&](const CUtensorMap* mh, const CUtensorMap* ml) {
                for (int kb = 0; kb < nkb; ++kb)
                    for (int mb = 0; mb < p.num_m; ++mb) {
                        mbar_wait<32>(a_empty(s), ph ^ 1);
                        const uint32_t st = a_base + s * A_PIECE;
                        mbar_arrive_expect_tx(a_full(s), A_PIECE);
                        tma_load_2d(mh, st, a_full(s), kb * BK, mb * BM);
                        tma_load_2d(ml, st + A_TILE, a_full(s), kb * BK, mb * BM);
                        if (++s == A_STAGES) { s = 0; ph ^= 1; }
                    }
            };
            for (uint32_t it = 0; it < n_my; ++it) {
                pieces(&map_a0_hi, &map_a0_lo);
                pieces(&map_a1_hi, &map_a1_lo);
            }
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer
        // Issue order per tile: G1(i), G2(i).  (Putting G1(i+1) in front of G2(i) measured slower: G2(i) then waits for
        // the whole transform of tile i+1, and E1(i+1) cannot write B2 before G2(i) has read it.)
        int sb = 0, sa = 0;
        uint32_t phb = 0, pha = 0;
        auto issue_kblock = [&](uint32_t bst, uint32_t d_base, int kb) {   // all m-blocks of one k-block
            for (int mb = 0; mb < p.num_m; ++mb) {
                mbar_wait(a_full(sa), pha);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t ast = a_base + sa * A_PIECE;
                    const uint32_t d_big = d_base + mb * 2 * BN;
#pragma unroll
                    for (int j = 0; j < BK / 16; ++j) {
                        // A (K-major, SWIZZLE_64B) and B (MN-major, SWIZZLE_128B panels) descriptors as in gemm_h.cu
                        const uint64_t a_hi = make_desc(ast + j * 32, 16, 512, 4);
                        const uint64_t a_lo = make_desc(ast + A_TILE + j * 32, 16, 512, 4);
                        const uint64_t b_hl = make_desc(bst + j * 2048, B_PANEL, 1024, 2);
                        umma_f16(d_big, a_hi, b_hl, G::IDESC_2N, (kb | j) != 0);
                        umma_f16(d_big + BN, a_lo, b_hl, G::IDESC_N, 1);
                    }
                    umma_commit(a_empty(sa));
                }
                __syncwarp();
                if (++sa == A_STAGES) { sa = 0; pha ^= 1; }
            }
        };
        auto g1 = [&](uint32_t it) {   // D1 = W0 * B(tile it); D1 is free once E1 of tile it-1 has read it
            mbar_wait(d1_empty, (it & 1) ^ 1);
            tc_fence_after();
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(b_ready(sb), phb);
                issue_kblock(b_base + sb * G::B_STAGE, tmem_base, kb);
                if (elect_one()) umma_commit(b_empty(sb));
                __syncwarp();
                if (++sb == B_STAGES) { sb = 0; phb ^= 1; }
            }
            if (elect_one()) umma_commit(d1_full);
            __syncwarp();
        };
        auto g2 = [&](uint32_t it) {   // D2 = W1 * B2(tile it); D2 is free once E2 of tile it-1 has read it
            mbar_wait(b2_ready, it & 1);
            mbar_wait(d2_empty, (it & 1) ^ 1);
            tc_fence_after();
            for (int kb = 0; kb < nkb; ++kb) issue_kblock(b2_base + kb * G::B_STAGE, tmem_base + 256, kb);
            if (elect_one()) umma_commit(d2_full);
            __syncwarp();
        };
        for (uint32_t it = 0; it < n_my; ++it) {
            g1(it);
            g2(it);
        }
    } else if (warp >= 8) {
        // ===================================================================== workers: transform of tile i, then E1 of tile i
        // Transform: raw fp32 tile -> B_hi / B_lo of the first GEMM.  E1: D1 -> depthwise-0 -> ELU -> split -> B2, two
        // warps per TMEM lane quarter (warp % 4), each owning half of the tile's columns, so E1 -- the longest serial
        // stage -- runs 8 warps wide while the epilogue warps drain the previous tile (E2).
        const int xw = warp - 8;
        const int q = warp & 3;                              // TMEM lane quarter this warp may read
        const int half = xw >> 2;                            // which half of the tile's columns
        constexpr int CH = BN / 64;                          // 32-column chunks per half and m-block
        const int row = q * 32 + lane;
        const f32x2 lo2 = pk2(1.0f / LO_SCALE, 1.0f / LO_SCALE);
        const float c_big0 = p.c_big0, c_inv0 = 1.0f / p.c_big0;
        float wk0h[5], bv0h = 0.f;
        auto load_taps = [&](int mb, float (&wk)[5], float& bv) {
            const int m = mb * BM + row;
            const bool ok = m < p.C;
#pragma unroll
            for (int k = 0; k < 5; ++k) wk[k] = ok ? __ldg(p.dw0_w + m * 5 + k) * c_big0 : 0.f;
            bv = (ok && p.dw0_b) ? __ldg(p.dw0_b + m) : 0.f;
        };
        if constexpr (NUM_M == 1 && RB_HOIST) load_taps(0, wk0h, bv0h);
        // the transform runs as two groups of 4 warps converting alternate k-blocks (8 k-rows per warp), so two
        // wait -> LDS -> math -> STS -> fence -> arrive chains are in flight (see gemm_h.cu)
        const int xg = xw / XW_PER_G, xl = xw % XW_PER_G;
        uint32_t n = 0;                                      // k-blocks seen by this CTA (all tiles)
        uint32_t it = 0;
        for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
            for (int kb = 0; kb < nkb; ++kb, ++n) {
                if ((int)(n % XG) != xg) continue;
                const int r = (int)(n % RAW_STAGES), s = (int)(n % B_STAGES);
                const uint32_t rph = (n / RAW_STAGES) & 1u, sph = (n / B_STAGES) & 1u;
                mbar_wait_ns(raw_full(r), rph, p.xform_sleep);
                mbar_wait_ns(b_empty(s), sph ^ 1, p.xform_sleep);
                const uint8_t* raw = gen_base + (raw_base - base) + r * G::RAW_BYTES;
                const uint32_t bdst = b_base + s * G::B_STAGE;
#pragma unroll
                for (int h = 0; h < XG; ++h) {
                    const int xq = xl * XG + h;              // 4-row slice of the box
                    if (p.pre == PRE_ELU) xform_tile_rows<PRE_ELU, BN>(raw, bdst, xq, lane, 1.0f);
                    else if (p.pre == PRE_SCALE_ELU) xform_tile_rows<PRE_SCALE_ELU, BN>(raw, bdst, xq, lane, p.pre_scale);
                    else xform_tile_rows<PRE_NONE, BN>(raw, bdst, xq, lane, 1.0f);
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(b_ready(s));
                    mbar_arrive(raw_empty(r));
                }
            }

            // ---------------------------------------------------------------- E1: D1 -> dw0 -> ELU -> split -> B2
            const int tt = (int)(tile % p.tiles_t);
            const int b = (int)(tile / p.tiles_t);
            const int tcol0 = tt * G::VAL - HALO;            // time of tile column 0
            const int n_chunks = min(BN / 32, (p.T - tcol0 + 31) / 32);
            const bool has_tail = tcol0 + BN > p.T - 4;      // the tile holds some of the last 4 time steps
            const int c_begin = half * CH, c_end = min(c_begin + CH, n_chunks);
            mbar_wait_ns(d2_full, (it & 1) ^ 1, p.xform_sleep);   // G2 of the previous tile has finished reading B2
            mbar_wait_ns(d1_full, it & 1, p.xform_sleep);
            tc_fence_after();
#pragma unroll 1
            for (int mb = 0; mb < NUM_M; ++mb) {
                const int mrow0 = mb * BM + q * 32;
                if (mrow0 >= p.C || c_begin >= c_end) continue;   // warp-uniform: weight-padding rows / nothing in this half
                const int m = mrow0 + lane;
                float wk[5], bv;
                if constexpr (NUM_M == 1 && RB_HOIST) {
#pragma unroll
                    for (int k = 0; k < 5; ++k) wk[k] = wk0h[k];
                    bv = bv0h;
                } else {
                    load_taps(mb, wk, bv);
                }
                const uint32_t t_d = tmem_base + ((uint32_t)(q * 32) << 16) + mb * 2 * BN;
                float carry[4] = {0.f, 0.f, 0.f, 0.f};
                if (c_begin > 0) {   // the 4 pointwise columns in front of this half
                    uint32_t cb[4], cs[4];
                    tmem_ld4(t_d + c_begin * 32 - 4, cb);
                    tmem_ld4(t_d + BN + c_begin * 32 - 4, cs);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 4; ++j) carry[j] = fmaf(__uint_as_float(cs[j]), 1.0f / LO_SCALE, __uint_as_float(cb[j]));
                }
                // row k = m of B2: k-block m / 32 = mb * 4 + q, row-in-block = lane
                const uint32_t b2row = b2_base + (uint32_t)(mb * 4 + q) * G::B_STAGE + (uint32_t)(lane >> 3) * 1024u +
                                       (uint32_t)(lane & 7) * 128u;
#pragma unroll 1
                for (int c = c_begin; c < c_end; ++c) {
                    uint32_t rb[32], rs[32];
                    tmem_ld32(t_d + c * 32, rb);
                    tmem_ld32(t_d + BN + c * 32, rs);
                    tmem_ld_wait();
                    float v[36];
#pragma unroll
                    for (int j = 0; j < 32; j += 2)
                        upk2(ffma2(pk2(__uint_as_float(rs[j]), __uint_as_float(rs[j + 1])), lo2,
                                   pk2(__uint_as_float(rb[j]), __uint_as_float(rb[j + 1]))), v[4 + j], v[5 + j]);
                    v[0] = carry[0]; v[1] = carry[1]; v[2] = carry[2]; v[3] = carry[3];
                    if (c == 0 && tt == 0) {   // tile columns 4..7 of the first tile are times -4..-1: the caller's cache
                        const float4 cv = *reinterpret_cast<const float4*>(p.c0_in + ((size_t)b * p.C + m) * 4);
                        v[8] = cv.x * c_inv0; v[9] = cv.y * c_inv0; v[10] = cv.z * c_inv0; v[11] = cv.w * c_inv0;
                    }
                    if (has_tail) {            // new cache = pointwise-0 outputs at times T-4..T-1 (owned by one tile each)
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int col = c * 32 + j;
                            const int t = tcol0 + col;
                            if (col >= HALO && t >= p.T - 4 && t < p.T)
                                p.c0_out[((size_t)b * p.C + m) * 4 + (t - (p.T - 4))] = v[4 + j] * c_big0;
                        }
                    }
#pragma unroll
                    for (int j8 = 0; j8 < 4; ++j8) {
                        float o[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const int i = j8 * 8 + e;   // output i of this chunk (tile column c*32 + i) uses v[i..i+4]
                            float a = bv;
#pragma unroll
                            for (int k = 0; k < 5; ++k) a = fmaf(wk[k], v[i + k], a);
                            o[e] = a;
                        }
                        uint32_t hi[4], lo[4];
                        elu_split8(o, hi, lo);
                        const uint32_t g8 = (uint32_t)(c * 4 + j8);   // 8-column group inside the tile
                        const uint32_t dst = b2row + (g8 >> 3) * B_PANEL + (((g8 & 7u) ^ (uint32_t)(lane & 7)) << 4);
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]),
                                     "r"(hi[3])
                                     : "memory");
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (uint32_t)G::B_TILE), "r"(lo[0]),
                                     "r"(lo[1]), "r"(lo[2]), "r"(lo[3])
                                     : "memory");
                    }
                    carry[0] = v[32]; carry[1] = v[33]; carry[2] = v[34]; carry[3] = v[35];
                }
            }
            tc_fence_before();
            mbar_arrive(d1_empty);
            fence_proxy_async();
            mbar_arrive(b2_ready);
        }
    } else if (warp >= 4 && warp < 8) {
        // ===================================================================== epilogue E2: D2 -> dw1 -> h += (TMA reduce-add)
        const int q = warp - 4;
        const int row = q * 32 + lane;                      // row inside an m-block = TMEM lane
        // TMA stores / waits of the epilogue: warp q == 0's elected lane (elect.sync picks the same lane every time for the
        // same mask, so the thread that commits a bulk group is the one that waits for it)
        [[maybe_unused]] const bool issuer = (q == 0 && lane == 0);
        const uint32_t sw = (uint32_t)(row & 7);
        const f32x2 lo2 = pk2(1.0f / LO_SCALE, 1.0f / LO_SCALE);
        const float c_big1 = p.c_big1, c_inv1 = 1.0f / p.c_big1;
        float wk1h[5], bv1h = 0.f;
        auto load_taps = [&](int mb, float (&wk)[5], float& bv) {
            const int m = mb * BM + row;
            const bool ok = m < p.C;
#pragma unroll
            for (int k = 0; k < 5; ++k) wk[k] = ok ? __ldg(p.dw1_w + m * 5 + k) * c_big1 : 0.f;
            bv = (ok && p.dw1_b) ? __ldg(p.dw1_b + m) : 0.f;
        };
        if constexpr (NUM_M == 1 && RB_HOIST) load_taps(0, wk1h, bv1h);
        uint32_t it = 0;
        uint32_t g = 0;                                      // running staging-chunk counter -> buffer parity
        for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
            const int tt = (int)(tile % p.tiles_t);
            const int b = (int)(tile / p.tiles_t);
            const int tcol0 = tt * G::VAL - HALO;
            const int n_chunks = min(BN / 32, (p.T - tcol0 + 31) / 32);
            const bool has_tail = tcol0 + BN > p.T - 4;
            mbar_wait<32>(d2_full, it & 1);
            tc_fence_after();
#pragma unroll 1
            for (int mb = 0; mb < NUM_M; ++mb) {
                const int mrow0 = mb * BM + q * 32;
                const bool warp_ok = mrow0 < p.C;
                if (mb * BM >= p.C) continue;                // whole m-block absent (uniform over the epilogue group)
                const int m = mrow0 + lane;
                float wk[5], bv;
                if constexpr (NUM_M == 1 && RB_HOIST) {
#pragma unroll
                    for (int k = 0; k < 5; ++k) wk[k] = wk1h[k];
                    bv = bv1h;
                } else {
                    load_taps(mb, wk, bv);
                }
                float carry[4] = {0.f, 0.f, 0.f, 0.f};
                const uint32_t t_d = tmem_base + ((uint32_t)(q * 32) << 16) + 256 + mb * 2 * BN;
#pragma unroll 1
                for (int c = 0; c < n_chunks; ++c, ++g) {
                    const uint32_t obuf = out_base + (g & 1) * OUT_BYTES;
                    float v[36];
                    if (warp_ok) {
                        uint32_t rb[32], rs[32];
                        tmem_ld32(t_d + c * 32, rb);
                        tmem_ld32(t_d + BN + c * 32, rs);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; j += 2)
                            upk2(ffma2(pk2(__uint_as_float(rs[j]), __uint_as_float(rs[j + 1])), lo2,
                                       pk2(__uint_as_float(rb[j]), __uint_as_float(rb[j + 1]))), v[4 + j], v[5 + j]);
                        v[0] = carry[0]; v[1] = carry[1]; v[2] = carry[2]; v[3] = carry[3];
                        if (c == 0 && tt == 0) {
                            const float4 cv = *reinterpret_cast<const float4*>(p.c1_in + ((size_t)b * p.C + m) * 4);
                            v[8] = cv.x * c_inv1; v[9] = cv.y * c_inv1; v[10] = cv.z * c_inv1; v[11] = cv.w * c_inv1;
                        }
                        if (has_tail) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                const int col = c * 32 + j;
                                const int t = tcol0 + col;
                                if (col >= HALO && t >= p.T - 4 && t < p.T)
                                    p.c1_out[((size_t)b * p.C + m) * 4 + (t - (p.T - 4))] = v[4 + j] * c_big1;
                            }
                        }
                    }
                    if (mb == p.num_m - 1 && c == n_chunks - 1) {   // last TMEM read of this tile: D2 may be overwritten
                        tc_fence_before();
                        mbar_arrive(d2_empty);
                    }
                    if (q == 0 && elect_one()) tma_wait_read<1>();   // the store that used this buffer two chunks ago has drained it
                    epi_bar_sync();
                    if (warp_ok) {
#pragma unroll
                        for (int j4 = 0; j4 < 8; ++j4) {
                            if (c == 0 && j4 < 2) continue;   // tile columns 0..7 are halo: they belong to the left neighbour
                            float o[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const int i = j4 * 4 + e;
                                float a = bv;
#pragma unroll
                                for (int k = 0; k < 5; ++k) a = fmaf(wk[k], v[i + k], a);
                                o[e] = a;
                            }
                            const uint32_t dst = c == 0 ? obuf + row * 96 + (j4 - 2) * 16 : obuf + row * 128 + (((uint32_t)j4 ^ sw) << 4);
                            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "f"(o[0]), "f"(o[1]), "f"(o[2]),
                                         "f"(o[3])
                                         : "memory");
                        }
                        carry[0] = v[32]; carry[1] = v[33]; carry[2] = v[34]; carry[3] = v[35];
                    }
                    fence_proxy_async();
                    epi_bar_sync();
                    if (q == 0 && elect_one()) {
                        if (c == 0) tma_reduce_add_3d(&map_y24, obuf, tcol0 + HALO, mb * BM, b);
                        else tma_reduce_add_3d(&map_y, obuf, tcol0 + c * 32, mb * BM, b);
                        tma_commit();
                    }
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

// The 8 columns in front of every tile, copied out before the in-place update: halo[b][c][tt * 8 + j] =
// h[b][c][tt * VAL - 8 + j] (zeros for the first tile, whose history comes from the caches).
__global__ void halo_gather_kernel(const float* __restrict__ h, long long bs, int rs, float* __restrict__ halo, int C,
                                   int tiles_t, int val) {
    const int c = blockIdx.y, b = blockIdx.z;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // (tile, half)
    if (idx >= tiles_t * 2) return;
    const int tt = idx >> 1, half = idx & 1;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tt > 0) v = *reinterpret_cast<const float4*>(h + b * bs + (long long)c * rs + (long long)tt * val - HALO + half * 4);
    *reinterpret_cast<float4*>(halo + (((size_t)b * C + c) * tiles_t + tt) * HALO + half * 4) = v;
}

}  // namespace rb

// ------------------------------------------------------------------------------- host side
// 128 < C <= 256: two 128-row blocks per 64-column tile (HILCODEC_RB_WIDE=1; measured slower than two fused-DWS launches in
// round 1, re-measured after the issuer work of round 2)
static bool rb_wide_on() {
    static const bool on = []() { const char* e = std::getenv("HILCODEC_RB_WIDE"); return e && e[0] == '1'; }();
    return on;
}
static int rb_bn(int C) { return C > 128 ? 64 : 128; }

bool resblock_h_usable(const PackedMat& W0, const PackedMat& W1, const float* h, long long bs, int rs, int T) {
    const int C = W0.M;
    if (W0.K != C || W1.M != C || W1.K != C) return false;
    if (!W0.H_hi || !W0.H_lo || !W1.H_hi || !W1.H_lo) return false;
    if (C < 32 || C > (rb_wide_on() ? 256 : 128) || (C & 31)) return false;
    if (T < 128) return false;                    // short chunks (streaming) keep the two-kernel path
    if ((rs & 3) || (bs & 3) || (reinterpret_cast<uintptr_t>(h) & 15)) return false;
    return true;
}

size_t resblock_h_halo_floats(int C, int T, int B) {
    const int val = rb_bn(C) - rb::HALO;
    return (size_t)B * C * ((T + val - 1) / val) * rb::HALO;
}

cudaError_t launch_resblock_halo(const float* h, long long bs, int rs, int B, int C, int T, float* halo, cudaStream_t st) {
    const int val = rb_bn(C) - rb::HALO;
    const int tiles_t = (T + val - 1) / val;
    if (C > 65535 || B > 65535) return cudaErrorInvalidValue;
    dim3 grid((tiles_t * 2 + 127) / 128, C, B);
    rb::halo_gather_kernel<<<grid, 128, 0, st>>>(h, bs, rs, halo, C, tiles_t, val);
    return cudaGetLastError();
}

template <int BN>
static cudaError_t launch_rb(const PackedMat& W0, const PackedMat& W1, float* h, long long bs, int rs, int B, int T, int pre,
                             float pre_scale, const float* dw0_w, const float* dw0_b, const float* dw1_w, const float* dw1_b,
                             const float* c0_in, float* c0_out, const float* c1_in, float* c1_out, const float* halo,
                             cudaStream_t st) {
    using namespace rb;
    using G = Geo<BN>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(resblock_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const int C = W0.M;
    const int tiles_t = (T + G::VAL - 1) / G::VAL;
    CUtensorMap a0h, a0l, a1h, a1l, mx, mh, my, my24;
    {
        const cuuint64_t dims[2] = {(cuuint64_t)W0.Kp32, (cuuint64_t)W0.Mp128};
        const cuuint64_t strides[1] = {(cuuint64_t)W0.Kp32 * 2};
        const cuuint32_t box[2] = {BK, BM};
        if (!tc::make_map_dt(&a0h, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, W0.H_hi, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B) ||
            !tc::make_map_dt(&a0l, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, W0.H_lo, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B) ||
            !tc::make_map_dt(&a1h, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, W1.H_hi, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B) ||
            !tc::make_map_dt(&a1l, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, W1.H_lo, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B))
            return cudaErrorInvalidValue;
    }
    {
        const cuuint64_t dims[3] = {(cuuint64_t)T, (cuuint64_t)C, (cuuint64_t)B};
        const cuuint64_t strides[2] = {(cuuint64_t)rs * 4, (cuuint64_t)bs * 4};
        const cuuint32_t box_x[3] = {(cuuint32_t)G::VAL, BK, 1};
        const cuuint32_t box_y[3] = {32, BM, 1};
        const cuuint32_t box_y24[3] = {24, BM, 1};
        if (!tc::make_map(&mx, h, 3, dims, strides, box_x, CU_TENSOR_MAP_SWIZZLE_NONE) ||
            !tc::make_map(&my, h, 3, dims, strides, box_y, CU_TENSOR_MAP_SWIZZLE_128B) ||
            !tc::make_map(&my24, h, 3, dims, strides, box_y24, CU_TENSOR_MAP_SWIZZLE_NONE))
            return cudaErrorInvalidValue;
    }
    {
        const cuuint64_t dims[3] = {(cuuint64_t)tiles_t * HALO, (cuuint64_t)C, (cuuint64_t)B};
        const cuuint64_t strides[2] = {(cuuint64_t)tiles_t * HALO * 4, (cuuint64_t)C * tiles_t * HALO * 4};
        const cuuint32_t box[3] = {HALO, BK, 1};
        if (!tc::make_map(&mh, halo, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return cudaErrorInvalidValue;
    }
    Params p{};
    p.C = C; p.T = T; p.B = B;
    p.nkb = C / BK;
    p.num_m = (C + BM - 1) / BM;
    p.tiles_t = tiles_t;
    p.total_tiles = (long long)tiles_t * B;
    p.pre = pre; p.pre_scale = (pre == PRE_SCALE_ELU) ? pre_scale : 1.0f;
    p.c_big0 = W0.h_inv_scale; p.c_big1 = W1.h_inv_scale;
    p.xform_sleep = tc::xform_sleep_env();
    p.dw0_w = dw0_w; p.dw0_b = dw0_b; p.dw1_w = dw1_w; p.dw1_b = dw1_b;
    p.c0_in = c0_in; p.c0_out = c0_out; p.c1_in = c1_in; p.c1_out = c1_out;
    const int num_sms = tc::device_sm_count();
    const unsigned grid = (unsigned)(p.total_tiles < num_sms ? p.total_tiles : num_sms);
    resblock_kernel<BN><<<grid, NUM_THREADS, G::SMEM, st>>>(a0h, a0l, a1h, a1l, mx, mh, my, my24, p);
    return cudaGetLastError();
}

// h is updated in place; `halo` must hold resblock_h_halo_floats(C, T, B) floats filled by launch_resblock_halo
// on the same stream.
cudaError_t launch_resblock_h(const PackedMat& W0, const PackedMat& W1, float* h, long long bs, int rs, int B, int T, int pre,
                              float pre_scale, const float* dw0_w, const float* dw0_b, const float* dw1_w, const float* dw1_b,
                              const float* c0_in, float* c0_out, const float* c1_in, float* c1_out, const float* halo,
                              cudaStream_t st) {
    if (B == 0 || T == 0) return cudaSuccess;
    if (W0.M > 128)
        return launch_rb<64>(W0, W1, h, bs, rs, B, T, pre, pre_scale, dw0_w, dw0_b, dw1_w, dw1_b, c0_in, c0_out, c1_in, c1_out,
                             halo, st);
    return launch_rb<128>(W0, W1, h, bs, rs, B, T, pre, pre_scale, dw0_w, dw0_b, dw1_w, dw1_b, c0_in, c0_out, c1_in, c1_out,
                          halo, st);
}

}  // namespace hil
