// Tensor-core pointwise GEMM with fp32-level accuracy from fp16 hi/lo splits (tcgen05 kind::f16).
//
//   Y[b][m][t] = sum_k W[m][k] * pre(X[b][k][t]) (+bias[m]) (+R[b][m][t])          (launch_gemm_h)
//   Y = dw5(W * pre(X)) + b_dw (+ skip), the whole DWSBlock (streaming.py:189-192)   (launch_gemm_h_dw)
//
// Same arithmetic idea as gemm_tc.cu (3 MMAs per product, lo*lo dropped) but every operand is a
// pair of fp16 numbers instead of a pair of tf32 numbers: an fp16 significand has the same 11 bits
// as tf32, the MMA runs at twice the tf32 rate and reads half the shared-memory bytes.
//   weights      w * 2^s = a_hi + a_lo * 2^-11   (s per matrix, so that max|w * 2^s| is in [2^13, 2^14);
//                                                 split once at hil_model_finalize)
//   activations  x       = b_hi + b_lo * 2^-11   (b_hi = fp16_rn(x), b_lo = fp16_rn((x - b_hi) * 2^11);
//                                                 requires |x| < 65504, otherwise the result is NaN)
//   big   += a_hi * b_hi                      (TMEM accumulator 0, fp32)
//   small += a_hi * b_lo + a_lo * b_hi        (TMEM accumulator 1, fp32, carries the 2^11 scale)
//   y      = (big + small * 2^-11) * 2^-s
// The 2^11 scale keeps both lo parts in the normal fp16 range, so nothing is lost to fp16
// subnormals; all scalings are powers of two (exact).
//
// Pipeline (one persistent CTA per SM, 128 x 128 output tile, BK = 32, 512 threads), three rings of 16 KB stages:
//   raw ring     (4): fp32 activation boxes [32 k][128 t] straight from TMA (no swizzle)
//   weight ring  (4): A_hi | A_lo (fp16, K-major, SWIZZLE_64B, TMA from the pre-split weights) -- or, for K <= 192, the
//                     CTA's weight rows resident for the whole kernel (Params::a_res)
//   operand ring (4): B_hi | B_lo (fp16, MN-major, SWIZZLE_128B, written by the transform warps)
// warp 0  X producer (TMA)      warp 3  A producer (TMA)      warp 1  MMA issuer
// warp 2  TMEM allocator        warps 4-7 epilogue            warps 8-15 transform (ELU prologue + split), two groups
// Per k-block two k16 steps of { A_hi x [B_hi | B_lo] (N = 256) -> [big | small]; A_lo x B_hi (N = 128) ->
// small }.  The MMA-issuing thread is the resource everything else waits for (HILCODEC_TRACE=1),
// so its serial path is kept minimal: ONE ready barrier per stage (weight bytes + transform arrivals), ONE commit, the
// next stage probed with mbarrier.test_wait before the current MMAs go out, and every single-thread region guarded by
// elect.sync (a generic `if (lane == 0)` makes ptxas wrap each tcgen05 / TMA instruction in an elect-and-loop waterfall).
// Epilogues are the ones of gemm_tc.cu (TMEM -> registers -> swizzled staging -> TMA store / reduce-add; optional fused
// causal depthwise k5, fused strided depthwise, fused transposed-depthwise prologue).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_fp16.h>

#include "h_split.cuh"

namespace hil {
namespace th {

using namespace tc;

constexpr int BM = 128, BN = 128, BK = 32;
constexpr int XG = 2;                             // k-blocks the transform warps convert concurrently (two groups of 4 warps)
constexpr int OUT_BUFS = 2;                       // epilogue staging buffers
// Three rings of 16 KB stages.  Every depth is a multiple of XG: a stage is then always filled for, and released by,
// the same transform group -- the groups wait by mbarrier parity, which is ambiguous for a stage shared by two groups
// (TMA loads complete out of order).  Stage s of the weight ring and stage s of the operand ring are filled for the same
// k-block and released together.
constexpr int RAW_STAGES = 4;                     // fp32 activation boxes from TMA
constexpr int A_STAGES = 4;                       // weight tiles A_hi | A_lo
constexpr int OP_STAGES = 4;                      // converted activations B_hi | B_lo
static_assert(A_STAGES == OP_STAGES, "stage s of both rings is one unit: one ready barrier, one release");
constexpr int RAW_BYTES = BK * BN * 4;            // 16 KB fp32 activation box
constexpr int A_TILE = BM * BK * 2;               // 8 KB
constexpr int B_TILE = BK * BN * 2;               // 8 KB
constexpr int A_STAGE = 2 * A_TILE;               // 16 KB: A_hi | A_lo
constexpr int OP_BYTES = 2 * B_TILE;              // 16 KB: B_hi panels | B_lo panels
constexpr int B_PANEL = 8 * 128 * (BK / 8);       // 4 KB: 64 columns x 32 k (4 swizzle atoms of 8 k-rows x 128 B)
constexpr int NUM_THREADS = 512;
constexpr int NUM_XFORM_WARPS = 8;
constexpr int NUM_EPI = 128;
constexpr int OUT_BYTES = BM * 32 * 4;            // 16 KB: one 128-row x 32-column output chunk
constexpr int TMEM_COLS = 512;
constexpr size_t SMEM_BYTES = 1024 + (size_t)RAW_STAGES * RAW_BYTES + (size_t)A_STAGES * A_STAGE + (size_t)OP_STAGES * OP_BYTES +
                              OUT_BUFS * OUT_BYTES + 256;   // 224 KB + slack

struct Params {
    int M, K, T, B;
    int num_m, tiles_t;
    long long total_tiles;
    int pre;
    int elu_poly;           // 1: elu_fast in the transform (HILCODEC_ELU_POLY=1), 0: ex2-only ELU on the packed pipe
    float pre_scale;
    float c_big, c_small;   // 2^-s and 2^-s * 2^-11
    const float* bias;
    int reduce_add;         // 1: Y += tile (TMA reduce), 0: Y = tile
    int xform_sleep;        // ns of back-off in the transform warps' barrier polls (0 = spin)
    // resident weights (K <= 192): the CTA's weight rows (all k-blocks, hi + lo) are loaded ONCE and stay in shared
    // memory for every tile it computes -- the grid is a multiple of num_m, so a CTA never changes its row block.  The
    // kernel is otherwise bound by the L2 -> SM operand stream (16 KB of weights + 16 KB of activations per k-block);
    // this halves it.  The tiles live in the four stages of the weight ring (k-blocks 0 .. 3) and, for K > 128, in the
    // top two stages of the operand ring (k-blocks 4, 5), which then has `nop` = 2 stages, one per transform group.
    int a_res, nraw, nop;
    int flat_t, flat_b;     // flat tiles (tc_flat_ok): samples per clip and number of clips; T = flat_b * flat_t columns, B = 1
    int probe;              // MMA issuer probes the next stage's barrier early (HILCODEC_MMA_PROBE=0: A/B off)
    // HILCODEC_TRACE=1 (tools/gpu/trace_gemm.py): per-CTA cycle counters of where each warp role waits, 24 per CTA:
    // 0 MMA<-tempty 1 MMA<-a_full 2 MMA<-b_ready 3 MMA total | 4 xform<-raw_full 5 xform<-op_empty 6 xform total |
    // 7 epi<-tfull 8 epi total 9 epi<-store drain | 10 Xprod<-raw_empty 11 Xprod total | 12 Aprod<-a_empty | 13 tiles |
    // 14 epi tcgen05.ld + wait 15 epi depthwise taps + staging stores (fused DWS epilogue only) | 16 MMA issue 17 commits
    unsigned long long* trace;
    int post_elu;           // fused DWS only: store ELU(y) (the consumer then needs no activation prologue)
    int t_step, t_halo;     // tile tt covers columns [tt * t_step - t_halo, ... + BN)
    const float* dw_w;      // [M][5]
    const float* dw_b;      // [M] or null
    const float* cache_in;  // [B][M][4]
    float* cache_out;       // [B][M][4]
    // kUp > 0: the B operand is the transposed depthwise conv of a low-rate tensor, computed in the transform
    int t_in;               // low-rate length (T = kUp * t_in)
    const float* up_w;      // [K][2 * kUp]
    const float* up_ci;     // [B][K]: activated last input of the previous chunk
    float* up_co;           // [B][K]
};

// columns of the low-rate input a 128-column output tile needs (one before its first column) + alignment slack
#ifndef HIL_UP_NI_ALIGN
#define HIL_UP_NI_ALIGN 8
#endif
// first low-rate column of tile tt's box: one before the tile's first input, rounded down to a 16-byte boundary
// (every TMA load in this library starts on one; an unaligned start raised "illegal instruction").  up_ni's
// padding (>= 3 columns beyond BN / S + 2) covers the shift.
template <int S>
__device__ __forceinline__ int up_box_start(int tt) { return ((tt * BN) / S - 1) & ~3; }
constexpr int up_ni(int S) { return ((BN / S + 2 + (BN % S ? 1 : 0)) + HIL_UP_NI_ALIGN - 1) & ~(HIL_UP_NI_ALIGN - 1); }

// Fused encoder downsampling (kDs = stride r > 0): 1x1 conv -> causal STRIDED depthwise conv (kernel 2r, stride r), the
// pair at the end of every encoder stage (streaming.py:506-510).  A tile's 128 pointwise columns are
// [unused alignment padding | r columns of history | DS_STEP new columns]; every TMA box starts on a 16-byte boundary
// (DS_HALO is a multiple of 4), tiles start on output boundaries (DS_STEP % r == 0) and write a multiple of 4 outputs
// (16-byte rows for the TMA store): r = 2 / 4 / 5 -> 60 / 28 / 24 outputs from 124 / 116 / 128 of the 128 columns.
// r = 8 (12 outputs from 104 columns, and a tensor-bound layer) stays on the two-kernel path.
constexpr int ds_halo(int r) { return (r + 3) & ~3; }
constexpr int ds_step(int r) { return ((BN - ds_halo(r)) / (4 * r)) * 4 * r; }
constexpr int ds_nout(int r) { return ds_step(r) / r; }
constexpr int ds_n1(int r) { return ((ds_nout(r) / 2 + 3) / 4) * 4; }   // outputs in the first of the two store halves
static_assert(ds_halo(2) + ds_step(2) == 124 && ds_halo(4) + ds_step(4) == 116 && ds_halo(5) + ds_step(5) == 128, "tile geometry");

constexpr uint32_t IDESC_N128 = make_idesc_f16(BM, BN);
constexpr uint32_t IDESC_N256 = make_idesc_f16(BM, 2 * BN);

// rows 4*xw .. 4*xw+3 of a [32 k][128 t] box -> B_hi / B_lo (MN-major, SWIZZLE_128B atoms of 8 k-rows x 128 B)
template <int kPre, bool kPoly = false>
__device__ __forceinline__ void xform_rows(const float4 (&v)[4], float s, int xw, uint32_t bhi, uint32_t chunk, uint32_t half8) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t k = (uint32_t)(xw * 4 + q);
        uint32_t h01, h23, l01, l23;
        split4<kPre, kPoly>(v[q], s, h01, h23, l01, l23);
        const uint32_t dst = bhi + (k >> 3) * 1024u + (k & 7u) * 128u + ((chunk ^ (k & 7u)) << 4) + half8;
        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(dst), "r"(h01), "r"(h23) : "memory");
        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(dst + (uint32_t)B_TILE), "r"(l01), "r"(l23) : "memory");
    }
}

// Upsample variant: B[k][n] = CausalConvTranspose1d(act(x))[k][n] (causal_layers.py:183-188, depthwise, kernel 2S,
// stride S) evaluated on the fly from the low-rate box x[k][i_start ..]:
//     B[k][n] = e[n/S] * w[k][n%S] + e[n/S - 1] * w[k][n%S + S],   e[i] = act(x[k][i]),  e[-1] = cache
// (two rounded products and one rounded add, like dwconvT_kernel), then the fp16 hi/lo split.  A lane owns 4
// consecutive output columns of 4 k-rows, which touch x[i0-1], x[i0] and -- unless S is a multiple of 4 -- x[i0+1].
template <int S, int kPre>
__device__ __forceinline__ void xform_rows_up(const float* raw, const float (&wreg)[4][8], int xw, int lane, uint32_t bhi,
                                              uint32_t chunk, uint32_t half8, float s, int n_abs0, int i_start, int k0, int K,
                                              const float* ci, float* co, int t_in) {
    constexpr int NI = up_ni(S);
    constexpr bool kNeedC = (S % 4) != 0;
    const int i0 = n_abs0 / S;
    const int r0 = n_abs0 - i0 * S;
    const int li = i0 - i_start;
    const int n_last = S * t_in - 1;                       // the column whose "current" input is x[t_in - 1]
    const bool owns_last = co != nullptr && n_abs0 <= n_last && n_last < n_abs0 + 4;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int kr = xw * 4 + q;
        const int k = k0 + kr;
        const float* xr = raw + kr * NI + li;
        float ea = xr[-1], eb = xr[0], ec = kNeedC ? xr[1] : 0.f;
        if (kPre != PRE_NONE) {
            eb = eb * s; ec = ec * s;
            const float tb = ex2_approx(eb * 1.4426950408889634f) - 1.0f;
            eb = eb > 0.f ? eb : tb;
            if (kNeedC) {
                const float tcv = ex2_approx(ec * 1.4426950408889634f) - 1.0f;
                ec = ec > 0.f ? ec : tcv;
            }
            // The input before this lane's first one is the previous lane's last one (S = 4: its `eb`, S = 2: its `ec`):
            // take its activated value by shuffle instead of evaluating the ELU a second time (same instructions on the
            // same number: identical bits); lane 0 takes it from the box.
            constexpr bool kShare = (S == 4 || S == 2);
            const float up = kShare ? __shfl_up_sync(0xffffffffu, S == 4 ? eb : ec, 1) : 0.f;
            if (kShare && lane > 0) {
                ea = up;
            } else {
                ea = ea * s;
                const float ta = ex2_approx(ea * 1.4426950408889634f) - 1.0f;
                ea = ea > 0.f ? ea : ta;
            }
        }
        if (i0 == 0) ea = k < K ? __ldg(ci + k) : 0.f;       // x[-1] is the cache (already activated)
        float u[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const bool next = kNeedC && (r0 + j >= S);       // this column already belongs to input i0 + 1
            const float cur = next ? ec : eb, prev = next ? eb : ea;
            u[j] = __fadd_rn(__fmul_rn(cur, wreg[q][j]), __fmul_rn(prev, wreg[q][4 + j]));
        }
        if (owns_last && k < K) co[k] = (n_last / S == i0) ? eb : ec;
        uint32_t h01, h23, l01, l23;
        split4<PRE_NONE>(make_float4(u[0], u[1], u[2], u[3]), 1.0f, h01, h23, l01, l23);
        const uint32_t kk = (uint32_t)kr;
        const uint32_t dst = bhi + (kk >> 3) * 1024u + (kk & 7u) * 128u + ((chunk ^ (kk & 7u)) << 4) + half8;
        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(dst), "r"(h01), "r"(h23) : "memory");
        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(dst + (uint32_t)B_TILE), "r"(l01), "r"(l23) : "memory");
    }
}

// taps of the 4 output columns n_abs0 .. n_abs0+3 for the 4 k-rows of this warp, from the k-block's tap slice
// [32][2S] that the producer copied into the raw stage behind the activation box (global loads here put an L2
// round trip on every k-block's critical path): wreg[q][j] = w[k][n%S], wreg[q][4+j] = w[k][n%S + S]
template <int S>
__device__ __forceinline__ void load_up_taps(const float* wsm, int xw, int n_abs0, float (&wreg)[4][8]) {
    const int r0 = n_abs0 % S;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float* wr = wsm + (xw * 4 + q) * 2 * S;
        if constexpr (S % 4 == 0) {
            const float4 a = *reinterpret_cast<const float4*>(wr + r0), b = *reinterpret_cast<const float4*>(wr + r0 + S);
            wreg[q][0] = a.x; wreg[q][1] = a.y; wreg[q][2] = a.z; wreg[q][3] = a.w;
            wreg[q][4] = b.x; wreg[q][5] = b.y; wreg[q][6] = b.z; wreg[q][7] = b.w;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int r = r0 + j;
                if (r >= S) r -= S;
                wreg[q][j] = wr[r];
                wreg[q][4 + j] = wr[r + S];
            }
        }
    }
}

// ------------------------------------------------------------------------------- kernel
// Variants measured and removed in round 2 (profiles/r2_ab_results.md): two epilogue groups draining alternate tiles
// (inside the box-to-box noise), fp16 hi/lo planes as the B operand with no transform pass (not faster: the operand
// stream out of L2 paces the wide layers, not the transform), activation boxes shared by TMA multicast across 2-CTA
// clusters (+5.7 % on the step: at cluster size 2 the multicast saves no L2 bandwidth and couples the two pipelines).
template <bool kDw, int kUp = 0, int kDs = 0>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_h_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
              const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_y,
              const __grid_constant__ CUtensorMap map_y28, const Params p) {
    constexpr int RAW_STAGES = th::RAW_STAGES;
    constexpr int NOUT = OUT_BUFS;                      // epilogue staging buffers
    constexpr int NUM_XW = NUM_XFORM_WARPS;             // transform warps
    constexpr int XW0 = NUM_THREADS / 32 - NUM_XW;      // first transform warp
    constexpr int XW_PER_G = NUM_XW / XG;               // warps per transform group (one k-block)
    constexpr int SPW = BK / 4 / XW_PER_G;              // 4-row slices of the box per transform warp
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t raw_base = base;
    const uint32_t a_base = raw_base + RAW_STAGES * RAW_BYTES;
    const uint32_t op_base = a_base + A_STAGES * A_STAGE;
    const uint32_t out_base = op_base + OP_STAGES * OP_BYTES;
    const uint32_t bars = out_base + NOUT * OUT_BYTES;
    // One `ready` barrier per operand stage collects the weight tile's TMA bytes AND the transform group's arrivals, and
    // one tcgen05.commit (`op_empty`) releases both halves: the single MMA-issuing thread is what paces this kernel
    // (HILCODEC_TRACE=1: ~180 cycles per mbarrier wait, ~100 per commit, ~70 per MMA on its serial path), so every
    // wait / commit it does not execute is time the tensor pipe gets.
    constexpr int NB0 = 2 * RAW_STAGES + 2 * OP_STAGES + 1;
    auto raw_full = [&](int r) { return bars + 8u * r; };
    auto raw_empty = [&](int r) { return bars + 8u * (RAW_STAGES + r); };
    auto ready = [&](int s) { return bars + 8u * (2 * RAW_STAGES + s); };
    auto op_empty = [&](int s) { return bars + 8u * (2 * RAW_STAGES + OP_STAGES + s); };
    const uint32_t ares_bar = bars + 8u * (2 * RAW_STAGES + 2 * OP_STAGES);   // resident weights loaded
    auto tfull_bar = [&](int a) { return bars + 8u * (NB0 + a); };
    auto tempty_bar = [&](int a) { return bars + 8u * (NB0 + 2 + a); };
    const uint32_t tmem_slot = bars + 8u * (NB0 + 4);
    uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb = (p.K + BK - 1) / BK;
    const int nraw = p.nraw, nop = p.nop;
    auto a_res_addr = [&](int kb) {   // resident weight tile of k-block kb (hi, then lo at + A_TILE)
        return kb < A_STAGES ? a_base + kb * A_STAGE : op_base + (OP_STAGES - 1 - (kb - A_STAGES)) * OP_BYTES;
    };
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_a_hi);
        prefetch_tmap(&map_a_lo);
        prefetch_tmap(&map_x);
        prefetch_tmap(&map_y);
    }
    if (warp == 1 && lane == 0) {
        for (int r = 0; r < RAW_STAGES; ++r) {
            mbar_init(raw_full(r), 1);
            mbar_init(raw_empty(r), XW_PER_G);
        }
        for (int s = 0; s < OP_STAGES; ++s) {
            mbar_init(ready(s), XW_PER_G + (p.a_res ? 0 : 1));   // transform warps (+ the weight producer's expect_tx)
            mbar_init(op_empty(s), 1);
        }
        mbar_init(ares_bar, 1);
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
        // ===================================================================== X producer (raw ring)
        if (elect_one()) {
            int r = 0;
            uint32_t ph = 0;
            long long tr_w = 0;
            const long long tr_s = p.trace ? clock64() : 0;
            for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const long long rest = tile / p.num_m;
                const int tt = (int)(rest % p.tiles_t);
                const int b = (int)(rest / p.tiles_t);
                for (int kb = 0; kb < nkb; ++kb) {
                    const long long tr0 = p.trace ? clock64() : 0;
                    mbar_wait<32>(raw_empty(r), ph ^ 1);
                    if (p.trace) tr_w += clock64() - tr0;
                    if constexpr (kUp > 0) {   // low-rate box: the inputs of the tile's columns and the one before
                        constexpr uint32_t XB = BK * up_ni(kUp) * 4, WB = BK * 2 * kUp * 4;   // activation box, tap slice
                        mbar_arrive_expect_tx(raw_full(r), XB + WB);
                        tma_load_3d(&map_x, raw_base + r * RAW_BYTES, raw_full(r), up_box_start<kUp>(tt), kb * BK, b);
                        bulk_load(raw_base + r * RAW_BYTES + XB, p.up_w + (size_t)kb * BK * 2 * kUp, WB, raw_full(r));
                    } else {
                        mbar_arrive_expect_tx(raw_full(r), RAW_BYTES);
                        if (p.flat_t)   // {t, clip, k}: the tile's 128 / T clips, all their samples
                            tma_load_3d(&map_x, raw_base + r * RAW_BYTES, raw_full(r), 0, tt * (BN / p.flat_t), kb * BK);
                        else
                            tma_load_3d(&map_x, raw_base + r * RAW_BYTES, raw_full(r), tt * p.t_step - p.t_halo, kb * BK, b);
                    }
                    if (++r == nraw) { r = 0; ph ^= 1; }
                }
            }
            if (p.trace) { p.trace[blockIdx.x * 24 + 10] = tr_w; p.trace[blockIdx.x * 24 + 11] = clock64() - tr_s; }
        }
    } else if (warp == 3) {
        // ===================================================================== A producer (op ring)
        const bool leader = elect_one();
        if (leader && p.a_res) {
            const int m_blk = (int)(blockIdx.x % p.num_m);   // == tile % num_m for every tile of this CTA
            mbar_arrive_expect_tx(ares_bar, (uint32_t)nkb * 2 * A_TILE);
            for (int kb = 0; kb < nkb; ++kb) {
                tma_load_2d(&map_a_hi, a_res_addr(kb), ares_bar, kb * BK, m_blk * BM);
                tma_load_2d(&map_a_lo, a_res_addr(kb) + A_TILE, ares_bar, kb * BK, m_blk * BM);
            }
        } else if (leader) {
            int s = 0;
            uint32_t ph = 0;
            for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const int m_blk = (int)(tile % p.num_m);
                for (int kb = 0; kb < nkb; ++kb) {
                    const long long tr0 = p.trace ? clock64() : 0;
                    mbar_wait<32>(op_empty(s), ph ^ 1);
                    if (p.trace) p.trace[blockIdx.x * 24 + 12] += clock64() - tr0;
                    const uint32_t st = a_base + s * A_STAGE;
                    mbar_arrive_expect_tx(ready(s), 2 * A_TILE);
                    tma_load_2d(&map_a_hi, st, ready(s), kb * BK, m_blk * BM);
                    tma_load_2d(&map_a_lo, st + A_TILE, ready(s), kb * BK, m_blk * BM);
                    if (++s == OP_STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer
        int s = 0;
        uint32_t ph = 0;
        long long it = 0;
        long long tr_e = 0, tr_a = 0, tr_b = 0, tr_i = 0, tr_c = 0;
        const long long tr_s = p.trace ? clock64() : 0;
        if (p.a_res) mbar_wait(ares_bar, 0);
        bool pre_ok = false;
        for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
            const int acc = (int)(it & 1);
            const uint32_t acc_ph = (uint32_t)((it >> 1) & 1);
            long long tr0 = p.trace ? clock64() : 0;
            mbar_wait(tempty_bar(acc), acc_ph ^ 1);
            if (p.trace) tr_e += clock64() - tr0;
            tc_fence_after();
            const uint32_t d_big = tmem_base + acc * 2 * BN;
            const uint32_t d_small = d_big + BN;
            for (int kb = 0; kb < nkb; ++kb) {
                const uint32_t st = op_base + s * OP_BYTES;
                const uint32_t at = p.a_res ? a_res_addr(kb) : a_base + s * A_STAGE;
                tr0 = p.trace ? clock64() : 0;
                if (!pre_ok) mbar_wait(ready(s), ph);
                if (p.trace) tr_b += clock64() - tr0;
                tc_fence_after();
                // probe the NEXT stage (non-blocking test_wait) before issuing this one's MMAs: the barrier round trip (~100 cycles on the
                // issuer's serial path) overlaps the issue below, and in steady state the transform is a stage ahead
                int s2 = s + 1;
                uint32_t ph2 = ph;
                if (s2 == nop) { s2 = 0; ph2 ^= 1; }
                pre_ok = p.probe && mbar_test_wait(ready(s2), ph2);
                if (elect_one()) {
                    const long long ti0 = p.trace ? clock64() : 0;
#pragma unroll
                    for (int j = 0; j < BK / 16; ++j) {
                        // A (K-major, SWIZZLE_64B): 8-row groups 512 B apart; +32 B walks one k16 step inside the
                        //   64-byte swizzle row.
                        // B (MN-major, SWIZZLE_128B): 64-column panels B_PANEL apart (LBO) -- B_hi p0, p1, B_lo p0,
                        //   p1 back to back -- 8-k-row atoms 1024 B apart (SBO); one k16 step = 2 atoms = 2 KB.
                        const uint64_t a_hi = make_desc(at + j * 32, 16, 512, 4);
                        const uint64_t a_lo = make_desc(at + A_TILE + j * 32, 16, 512, 4);
                        const uint64_t b_hl = make_desc(st + j * 2048, B_PANEL, 1024, 2);
                        umma_f16(d_big, a_hi, b_hl, IDESC_N256, (kb | j) != 0);
                        umma_f16(d_small, a_lo, b_hl, IDESC_N128, 1);
                    }
                    const long long ti1 = p.trace ? clock64() : 0;
                    umma_commit(op_empty(s));
                    if (kb == nkb - 1) umma_commit(tfull_bar(acc));
                    if (p.trace) { tr_i += ti1 - ti0; tr_c += clock64() - ti1; }
                }
                __syncwarp();
                s = s2; ph = ph2;
            }
        }
        if (p.trace && lane == 0) {
            unsigned long long* tr = p.trace + blockIdx.x * 24;
            tr[0] = tr_e; tr[1] = tr_a; tr[2] = tr_b; tr[3] = clock64() - tr_s; tr[13] = it; tr[16] = tr_i; tr[17] = tr_c;
        }
    } else if (warp >= XW0) {
        // ===================================================================== transform: raw fp32 -> B_hi / B_lo fp16
        // The transform warps work as XG groups, each group converting every XG-th k-block (a warp owns SPW 4-row
        // slices of the box; a lane owns 4 consecutive columns of a row: conflict-free LDS.128, STS.64 into the 128B-swizzled MN-major
        // atoms).  With one group the conversion of a k-block is one serial latency chain (wait -> LDS -> math -> STS
        // -> fence -> arrive, ~1000 cycles measured) that paces the whole mainloop; XG chains overlap.
        const int xw = warp - XW0;
        const int xg = xw / XW_PER_G, xl = xw % XW_PER_G;
        const uint32_t panel = (uint32_t)(lane >> 4);     // 64-column panel
        const uint32_t chunk = (uint32_t)((lane >> 1) & 7);  // 16-byte chunk (8 columns) inside the 128-byte row
        const uint32_t half8 = (uint32_t)(lane & 1) * 8u;
        uint32_t n = 0;                                    // k-blocks seen by this CTA (all tiles)
        // ring positions of this group's next k-block (n = xg, xg + XG, ...): advanced by XG with wrap-around, the
        // parity flips on every wrap (the depths are multiples of XG or, for the operand ring, at least XG)
        int r = xg % nraw, s = xg % nop;
        uint32_t rph = 0, sph = 0;
        long long tr_r = 0, tr_o = 0;
        const long long tr_s = p.trace ? clock64() : 0;
        for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            [[maybe_unused]] const int m_blk = (int)(tile % p.num_m);
            [[maybe_unused]] const long long rest = tile / p.num_m;
            [[maybe_unused]] const int tt = (int)(rest % p.tiles_t);
            [[maybe_unused]] const int b = (int)(rest / p.tiles_t);
            for (int kb = 0; kb < nkb; ++kb, ++n) {
                if ((int)(n % XG) != xg) continue;
                long long tr0 = p.trace ? clock64() : 0;
                mbar_wait_ns(raw_full(r), rph, p.xform_sleep);
                if (p.trace) { const long long t1 = clock64(); tr_r += t1 - tr0; tr0 = t1; }
                mbar_wait_ns(op_empty(s), sph ^ 1, p.xform_sleep);
                if (p.trace) tr_o += clock64() - tr0;
                const uint32_t bhi = op_base + s * OP_BYTES + panel * B_PANEL;
                if constexpr (kUp > 0) {
                    const float* raw = reinterpret_cast<const float*>(gen_base + (raw_base - base) + r * RAW_BYTES);
                    const float* ci = p.up_ci + (size_t)b * p.K;
                    float* co = m_blk == 0 ? p.up_co + (size_t)b * p.K : nullptr;
#pragma unroll
                    for (int h = 0; h < SPW; ++h) {
                        const int xq = xl * SPW + h;           // 4-row slice of the box
                        float wreg[4][8];
                        load_up_taps<kUp>(raw + BK * up_ni(kUp), xq, tt * BN + 4 * lane, wreg);
                        if (p.pre == PRE_NONE)
                            xform_rows_up<kUp, PRE_NONE>(raw, wreg, xq, lane, bhi, chunk, half8, 1.0f, tt * BN + 4 * lane,
                                                         up_box_start<kUp>(tt), kb * BK, p.K, ci, co, p.t_in);
                        else
                            xform_rows_up<kUp, PRE_SCALE_ELU>(raw, wreg, xq, lane, bhi, chunk, half8, p.pre_scale,
                                                              tt * BN + 4 * lane, up_box_start<kUp>(tt), kb * BK, p.K, ci, co,
                                                              p.t_in);
                    }
                } else {
                    const float4* src = reinterpret_cast<const float4*>(gen_base + (raw_base - base) + r * RAW_BYTES);
                    float4 v[SPW][4];
#pragma unroll
                    for (int h = 0; h < SPW; ++h)
#pragma unroll
                        for (int q = 0; q < 4; ++q) v[h][q] = src[((xl * SPW + h) * 4 + q) * (BN / 4) + lane];
#pragma unroll
                    for (int h = 0; h < SPW; ++h) {
                        const int xq = xl * SPW + h;
                        if (p.pre == PRE_NONE) xform_rows<PRE_NONE>(v[h], 1.0f, xq, bhi, chunk, half8);
                        else if (p.elu_poly) xform_rows<PRE_SCALE_ELU, true>(v[h], p.pre_scale, xq, bhi, chunk, half8);
                        else if (p.pre == PRE_ELU) xform_rows<PRE_ELU>(v[h], 1.0f, xq, bhi, chunk, half8);
                        else xform_rows<PRE_SCALE_ELU>(v[h], p.pre_scale, xq, bhi, chunk, half8);
                    }
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(ready(s));
                    mbar_arrive(raw_empty(r));
                }
                r += XG; if (r >= nraw) { r -= nraw; rph ^= 1u; }
                s += XG; if (s >= nop) { s -= nop; sph ^= 1u; }
            }
        }
        if (p.trace && xw == 0 && lane == 0) {
            unsigned long long* tr = p.trace + blockIdx.x * 24;
            tr[4] = tr_r; tr[5] = tr_o; tr[6] = clock64() - tr_s;
        }
    } else if (warp >= 4 && warp < 8) {
        // ===================================================================== epilogue
        const int q = warp & 3;                             // TMEM lane quarter of this warp
        const int row = q * 32 + lane;                      // row inside the 128-row tile = TMEM lane
        // TMA stores / waits of the epilogue: warp q == 0's elected lane (elect.sync picks the same lane every time for the
        // same mask, so the thread that commits a bulk group is the one that waits for it)
        [[maybe_unused]] const bool issuer = (q == 0 && lane == 0);
        const uint32_t my_out = out_base;
        auto epi_bar_sync = [&]() { asm volatile("bar.sync 1, 128;" ::: "memory"); };
        // 128B-swizzle phase of this row; flat tiles stage unswizzled (a SWIZZLE_128B tensor map whose box is only
        // 32 bytes wide faults on the store: tools/gpu/tma_flat_test.cu)
        const uint32_t sw = p.flat_t ? 0u : (uint32_t)(row & 7);
        const float c_big = p.c_big;                        // 2^-s; c_small = c_big * 2^-11
        const f32x2 lo2 = pk2(1.0f / LO_SCALE, 1.0f / LO_SCALE), cb2 = pk2(c_big, c_big);
        long long it = 0;
        uint32_t g = 0;                                      // running chunk counter -> staging buffer parity
        long long tr_f = 0, tr_d = 0, tr_l = 0, tr_c = 0;
        const long long tr_s = p.trace ? clock64() : 0;
        if constexpr (kDs > 0) {
            // ---- fused strided depthwise epilogue: y[n] = b + sum_{k < 2r} w[k] * pw[(n - 1) r + k], pw[-r .. -1] = cache.
            // A thread owns one channel row and walks the tile's columns once; input column j feeds output j / r with
            // tap (j % r) + r and output j / r + 1 with tap j % r, so two running sums see the taps in the order
            // k = 0 .. 2r - 1 of the stand-alone kernel (dwconv_strided4_kernel, conv.cu): bit-identical results.
            // The tile's NOUT outputs leave in two halves (N1 | N2, both multiples of 4) through the two staging buffers,
            // so the TMA store of one half drains while the other half is computed (a single store per tile, waited for
            // before the next tile may touch the buffer, made this epilogue 3x slower than the kernel pair it replaces).
            constexpr int R = kDs, HALO = ds_halo(R), STEP = ds_step(R), NOUT = ds_nout(R), J0 = HALO - R;
            constexpr int N1 = ds_n1(R), N2 = NOUT - N1;
            static_assert(N1 % 4 == 0 && N2 % 4 == 0 && N1 * 4 * BM <= OUT_BYTES && N2 * 4 * BM <= OUT_BYTES && OUT_BUFS >= 2, "staging");
            const uint32_t buf_a = my_out, buf_b = my_out + OUT_BYTES;
            float wk[2 * R], bv = 0.f;
#pragma unroll
            for (int k = 0; k < 2 * R; ++k) wk[k] = 0.f;
            int m_taps = -1;
            for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
                const int m_blk = (int)(tile % p.num_m);
                const long long rest = tile / p.num_m;
                const int tt = (int)(rest % p.tiles_t);
                const int b = (int)(rest / p.tiles_t);
                const int acc = (int)(it & 1);
                const uint32_t acc_ph = (uint32_t)((it >> 1) & 1);
                const int m = m_blk * BM + row;
                const bool row_ok = m < p.M;
                if (m != m_taps) {
#pragma unroll
                    for (int k = 0; k < 2 * R; ++k) wk[k] = row_ok ? p.dw_w[m * 2 * R + k] * c_big : 0.f;
                    bv = (row_ok && p.dw_b) ? p.dw_b[m] : 0.f;
                    m_taps = m;
                }
                const float c_inv = 1.0f / c_big;
                const int tcol0 = tt * STEP - HALO;               // time of tile column 0
                float hist[R];                                     // tile 0: the r pointwise outputs before the chunk
#pragma unroll
                for (int k = 0; k < R; ++k) hist[k] = (tt == 0 && row_ok) ? p.cache_in[((size_t)b * p.M + m) * R + k] * c_inv : 0.f;
                // the tile whose new columns hold times T-r .. T-1 (T % r == 0 and STEP % r == 0: all of them or none)
                const int j_tail = p.T - R - tcol0;
                const bool tail_tile = j_tail >= HALO && j_tail < HALO + STEP;
                mbar_wait<64>(tfull_bar(acc), acc_ph);
                tc_fence_after();
                const uint32_t t_big = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 2 * BN;
                if (q == 0 && elect_one()) tma_wait_read<0>();                    // the previous tile's stores have drained both buffers
                epi_bar_sync();
                float a_cur = 0.f, a_next = 0.f;
                float o4[4];
#pragma unroll
                for (int c = 0; c < BN / 32; ++c) {
                    uint32_t rb[32], rs[32];
                    tmem_ld32(t_big + c * 32, rb);
                    tmem_ld32(t_big + BN + c * 32, rs);
                    tmem_ld_wait();
                    if (c == BN / 32 - 1) {
                        if (tail_tile) {   // new cache = pointwise outputs at times T-r .. T-1, re-read from TMEM (one tile per clip)
                            // 8 columns from the 4-column boundary below j_tail (TMEM loads of this library always
                            // start on one), then a compile-time-indexed select of the r values
                            const int jb = j_tail & ~3, off = j_tail & 3;
                            uint32_t cb[8], cs[8];
                            {
                                uint32_t t0[4], t1[4], u0[4], u1[4];
                                tmem_ld4(t_big + jb, t0);
                                tmem_ld4(t_big + jb + 4, t1);
                                tmem_ld4(t_big + BN + jb, u0);
                                tmem_ld4(t_big + BN + jb + 4, u1);
                                tmem_ld_wait();
#pragma unroll
                                for (int k = 0; k < 4; ++k) { cb[k] = t0[k]; cb[4 + k] = t1[k]; cs[k] = u0[k]; cs[4 + k] = u1[k]; }
                            }
                            if (row_ok) {
#pragma unroll
                                for (int k = 0; k < R; ++k) {
                                    const uint32_t bg = off == 0 ? cb[k] : off == 1 ? cb[k + 1] : off == 2 ? cb[k + 2] : cb[k + 3];
                                    const uint32_t sm = off == 0 ? cs[k] : off == 1 ? cs[k + 1] : off == 2 ? cs[k + 2] : cs[k + 3];
                                    p.cache_out[((size_t)b * p.M + m) * R + k] =
                                        fmaf(__uint_as_float(sm), 1.0f / LO_SCALE, __uint_as_float(bg)) * c_big;
                                }
                            }
                        }
                        tc_fence_before();
                        mbar_arrive(tempty_bar(acc));
                    }
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; j += 2)
                        upk2(ffma2(pk2(__uint_as_float(rs[j]), __uint_as_float(rs[j + 1])), lo2,
                                   pk2(__uint_as_float(rb[j]), __uint_as_float(rb[j + 1]))), v[j], v[j + 1]);
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int j = c * 32 + i;                  // tile column (compile-time after unrolling)
                        if (j < J0 || j >= HALO + STEP) continue;
                        const int jj = j - J0;                     // 0 .. R + STEP - 1
                        const int grp = jj / R - 1, ph = jj % R;   // group -1 = history, 0 .. NOUT-1 = outputs
                        float x = v[i];
                        if (grp < 0 && tt == 0) x = hist[ph];
                        if (grp >= 0) a_cur = fmaf(wk[ph + R], x, a_cur);
                        if (grp < NOUT - 1) a_next = fmaf(wk[ph], x, a_next);
                        if (ph == R - 1) {
                            if (grp >= 0) {
                                o4[grp & 3] = a_cur + bv;
                                if ((grp & 3) == 3) {
                                    const uint32_t dst = grp < N1 ? buf_a + row * (N1 * 4) + (uint32_t)(grp - 3) * 4u
                                                                  : buf_b + row * (N2 * 4) + (uint32_t)(grp - N1 - 3) * 4u;
                                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "f"(o4[0]), "f"(o4[1]),
                                                 "f"(o4[2]), "f"(o4[3]) : "memory");
                                }
                                if (grp == N1 - 1) {               // first half complete: hand it to the TMA engine
                                    fence_proxy_async();
                                    epi_bar_sync();
                                    if (q == 0 && elect_one()) {
                                        tma_store_3d(&map_y, buf_a, tt * NOUT, m_blk * BM, b);
                                        tma_commit();
                                    }
                                }
                            }
                            a_cur = a_next;
                            a_next = 0.f;
                        }
                    }
                }
                fence_proxy_async();
                epi_bar_sync();
                if (q == 0 && elect_one()) {
                    tma_store_3d(&map_y28, buf_b, tt * NOUT + N1, m_blk * BM, b);
                    tma_commit();
                }
            }
            if (q == 0 && elect_one()) tma_wait_all();
        } else if constexpr (!kDw) {
            float bv = 0.f;
            int m_bias = -1;
            for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
                const int m_blk = (int)(tile % p.num_m);
                const long long rest = tile / p.num_m;
                const int tt = (int)(rest % p.tiles_t);
                const int b = (int)(rest / p.tiles_t);
                const int acc = (int)(it & 1);
                const uint32_t acc_ph = (uint32_t)((it >> 1) & 1);
                const long long tr0 = p.trace ? clock64() : 0;
                mbar_wait<64>(tfull_bar(acc), acc_ph);
                if (p.trace) tr_f += clock64() - tr0;
                tc_fence_after();
                const int m = m_blk * BM + row;
                if (m != m_bias) { bv = (m < p.M && p.bias) ? p.bias[m] : 0.f; m_bias = m; }
                const f32x2 bv2 = pk2(bv, bv);
                const int t0 = tt * BN;
                const int n_chunks = min(BN / 32, (p.T - t0 + 31) / 32);
                const uint32_t t_big = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 2 * BN;
#pragma unroll 1
                for (int c = 0; c < n_chunks; ++c, ++g) {
                    const uint32_t obuf = my_out + (g % NOUT) * OUT_BYTES;
                    const long long tr1 = p.trace ? clock64() : 0;
                    if (q == 0 && elect_one()) tma_wait_read<NOUT - 1>();          // the store that used this buffer two chunks ago has drained it
                    epi_bar_sync();
                    if (p.trace) tr_d += clock64() - tr1;
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
                        // (big + small * 2^-11) * 2^-s + bias: two packed FMAs per pair; bit-identical to
                        // fmaf(small, c_small, big * c_big) + bias because every scale is a power of two
                        float o[4];
                        upk2(ffma2(ffma2(pk2(__uint_as_float(rs[4 * j]), __uint_as_float(rs[4 * j + 1])), lo2,
                                         pk2(__uint_as_float(rb[4 * j]), __uint_as_float(rb[4 * j + 1]))), cb2, bv2), o[0], o[1]);
                        upk2(ffma2(ffma2(pk2(__uint_as_float(rs[4 * j + 2]), __uint_as_float(rs[4 * j + 3])), lo2,
                                         pk2(__uint_as_float(rb[4 * j + 2]), __uint_as_float(rb[4 * j + 3]))), cb2, bv2), o[2], o[3]);
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(orow + (((uint32_t)j ^ sw) << 4)), "f"(o[0]),
                                     "f"(o[1]), "f"(o[2]), "f"(o[3])
                                     : "memory");
                    }
                    fence_proxy_async();
                    epi_bar_sync();
                    if (q == 0 && elect_one()) {
                        // flat tiles: the chunk's 32 columns are 32 / T whole clips of the {t, clip, m} map
                        const int c0 = p.flat_t ? 0 : t0 + c * 32;
                        const int c1 = p.flat_t ? (t0 + c * 32) / p.flat_t : m_blk * BM;
                        const int c2 = p.flat_t ? m_blk * BM : b;
                        if (p.reduce_add) tma_reduce_add_3d(&map_y, obuf, c0, c1, c2);
                        else tma_store_3d(&map_y, obuf, c0, c1, c2);
                        tma_commit();
                    }
                }
            }
            if (q == 0 && elect_one()) tma_wait_all();
        } else {
            // ---- fused DWS epilogue (see gemm_tc.cu): the tile holds 128 pointwise columns for times
            // [t0-4, t0+124); each thread owns one channel row and slides the 5-tap window along it in registers.
            float wk[5] = {0.f, 0.f, 0.f, 0.f, 0.f}, bv = 0.f;
            int m_taps = -1;
            if (p.flat_t) {
                // ---- flat tiles (8 samples per clip: 16 whole clips per tile, 4 per 32-column chunk).  Every clip's window
                // starts from ITS cache (times -4 .. -1) and every column is an output, so there is no halo column and
                // no carry between chunks: out[t] = b + sum_k w[k] * v[t + k], v = [cache | 8 pointwise values].
                constexpr int TT = 8, CPC = 32 / TT;
                for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
                    const int m_blk = (int)(tile % p.num_m);
                    const int tt = (int)(tile / p.num_m);
                    const int acc = (int)(it & 1);
                    const uint32_t acc_ph = (uint32_t)((it >> 1) & 1);
                    const int m = m_blk * BM + row;
                    const bool row_ok = m < p.M;
                    if (m != m_taps) {
#pragma unroll
                        for (int k = 0; k < 5; ++k) wk[k] = row_ok ? p.dw_w[m * 5 + k] * c_big : 0.f;
                        bv = (row_ok && p.dw_b) ? p.dw_b[m] : 0.f;
                        m_taps = m;
                    }
                    const float c_inv = 1.0f / c_big;
                    const int clip0 = tt * (BN / TT);                 // first clip of the tile
                    const int n_chunks = min(BN / 32, (p.flat_b - clip0 + CPC - 1) / CPC);
                    mbar_wait<64>(tfull_bar(acc), acc_ph);
                    tc_fence_after();
                    const uint32_t t_big = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 2 * BN;
#pragma unroll 1
                    for (int c = 0; c < n_chunks; ++c, ++g) {
                        const uint32_t obuf = my_out + (g % NOUT) * OUT_BYTES;
                        float4 cv[CPC];                               // the 4 clips' caches: loads in flight under the TMEM read
#pragma unroll
                        for (int j = 0; j < CPC; ++j) {
                            const int clip = clip0 + c * CPC + j;
                            cv[j] = (row_ok && clip < p.flat_b)
                                        ? *reinterpret_cast<const float4*>(p.cache_in + ((size_t)clip * p.M + m) * 4)
                                        : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                        uint32_t rb[32], rs[32];
                        tmem_ld32(t_big + c * 32, rb);
                        tmem_ld32(t_big + BN + c * 32, rs);
                        tmem_ld_wait();
                        if (c == n_chunks - 1) {
                            tc_fence_before();
                            mbar_arrive(tempty_bar(acc));
                        }
                        if (q == 0 && elect_one()) tma_wait_read<NOUT - 1>();
                        epi_bar_sync();
#pragma unroll
                        for (int j = 0; j < CPC; ++j) {
                            float v[4 + TT];
                            v[0] = cv[j].x * c_inv; v[1] = cv[j].y * c_inv; v[2] = cv[j].z * c_inv; v[3] = cv[j].w * c_inv;
#pragma unroll
                            for (int i = 0; i < TT; i += 2)
                                upk2(ffma2(pk2(__uint_as_float(rs[j * TT + i]), __uint_as_float(rs[j * TT + i + 1])), lo2,
                                           pk2(__uint_as_float(rb[j * TT + i]), __uint_as_float(rb[j * TT + i + 1]))),
                                     v[4 + i], v[5 + i]);
                            const int clip = clip0 + c * CPC + j;
                            if (row_ok && clip < p.flat_b)   // new cache = the clip's last 4 pointwise values, true scale
                                *reinterpret_cast<float4*>(p.cache_out + ((size_t)clip * p.M + m) * 4) =
                                    make_float4(v[TT] * c_big, v[TT + 1] * c_big, v[TT + 2] * c_big, v[TT + 3] * c_big);
#pragma unroll
                            for (int h = 0; h < TT / 4; ++h) {
                                float o[4];
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    float a = bv;
#pragma unroll
                                    for (int k = 0; k < 5; ++k) a = fmaf(wk[k], v[h * 4 + e + k], a);
                                    o[e] = a;
                                }
                                if (p.post_elu) elu4(o[0], o[1], o[2], o[3]);
                                const uint32_t dst = obuf + row * 128 + ((((uint32_t)(j * (TT / 4) + h)) ^ sw) << 4);
                                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "f"(o[0]), "f"(o[1]), "f"(o[2]),
                                             "f"(o[3])
                                             : "memory");
                            }
                        }
                        fence_proxy_async();
                        epi_bar_sync();
                        if (q == 0 && elect_one()) {
                            if (p.reduce_add) tma_reduce_add_3d(&map_y, obuf, 0, clip0 + c * CPC, m_blk * BM);
                            else tma_store_3d(&map_y, obuf, 0, clip0 + c * CPC, m_blk * BM);
                            tma_commit();
                        }
                    }
                }
            } else
            for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
                const int m_blk = (int)(tile % p.num_m);
                const long long rest = tile / p.num_m;
                const int tt = (int)(rest % p.tiles_t);
                const int b = (int)(rest / p.tiles_t);
                const int acc = (int)(it & 1);
                const uint32_t acc_ph = (uint32_t)((it >> 1) & 1);
                const int m = m_blk * BM + row;
                const bool row_ok = m < p.M;
                // The window slides over u = (big + small * 2^-11) = pointwise / c_big, with the taps pre-multiplied
                // by c_big (powers of two: the products are unchanged); the caches hold true-scale values.
                if (m != m_taps) {   // taps and bias of this row: global loads on the epilogue's critical path, so only
                                     // when the row changes (never, when the CTA keeps its row block: resident weights)
#pragma unroll
                    for (int k = 0; k < 5; ++k) wk[k] = row_ok ? p.dw_w[m * 5 + k] * c_big : 0.f;
                    bv = (row_ok && p.dw_b) ? p.dw_b[m] : 0.f;
                    m_taps = m;
                }
                const float c_inv = 1.0f / c_big;
                const int tcol0 = tt * p.t_step - p.t_halo;       // time of tile column 0
                const int n_chunks = min(BN / 32, (p.T - tcol0 + 31) / 32);
                const bool has_tail = row_ok && (tcol0 + BN > p.T - 4);  // tile holds some of the last 4 columns
                float carry[4] = {0.f, 0.f, 0.f, 0.f};
                if (tt == 0 && row_ok) {
                    const float4 cv = *reinterpret_cast<const float4*>(p.cache_in + ((size_t)b * p.M + m) * 4);
                    carry[0] = cv.x * c_inv; carry[1] = cv.y * c_inv; carry[2] = cv.z * c_inv; carry[3] = cv.w * c_inv;
                }
                const long long tr0 = p.trace ? clock64() : 0;
                mbar_wait<64>(tfull_bar(acc), acc_ph);
                if (p.trace) tr_f += clock64() - tr0;
                tc_fence_after();
                const uint32_t t_big = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 2 * BN;
#pragma unroll 1
                for (int c = 0; c < n_chunks; ++c, ++g) {
                    const uint32_t obuf = my_out + (g % NOUT) * OUT_BYTES;
                    uint32_t rb[32], rs[32];
                    const long long tr2 = p.trace ? clock64() : 0;
                    tmem_ld32(t_big + c * 32, rb);
                    tmem_ld32(t_big + BN + c * 32, rs);
                    tmem_ld_wait();
                    if (p.trace) tr_l += clock64() - tr2;
                    if (c == n_chunks - 1) {
                        tc_fence_before();
                        mbar_arrive(tempty_bar(acc));
                    }
                    float v[36];
#pragma unroll
                    for (int j = 0; j < 32; j += 2)
                        upk2(ffma2(pk2(__uint_as_float(rs[j]), __uint_as_float(rs[j + 1])), lo2,
                                   pk2(__uint_as_float(rb[j]), __uint_as_float(rb[j + 1]))), v[4 + j], v[5 + j]);
                    // columns 0..3 of the first tile are times -4..-1: the cache handed in by the caller
                    if (c == 0 && tt == 0) { v[4] = carry[0]; v[5] = carry[1]; v[6] = carry[2]; v[7] = carry[3]; }
                    v[0] = carry[0]; v[1] = carry[1]; v[2] = carry[2]; v[3] = carry[3];
                    if (has_tail) {  // new cache = pointwise outputs at times T-4..T-1 (each time is owned by one tile)
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int col = c * 32 + j;
                            const int t = tcol0 + col;
                            if (col >= p.t_halo && t >= p.T - 4 && t < p.T)
                                p.cache_out[((size_t)b * p.M + m) * 4 + (t - (p.T - 4))] = v[4 + j] * c_big;
                        }
                    }
                    const long long tr1 = p.trace ? clock64() : 0;
                    if (q == 0 && elect_one()) tma_wait_read<NOUT - 1>();   // the store that used this buffer two chunks ago has drained it
                    epi_bar_sync();
                    if (p.trace) tr_d += clock64() - tr1;
                    const long long tr3 = p.trace ? clock64() : 0;
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {
                        float o[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int i = j4 * 4 + e;  // output i of this chunk uses v[i..i+4]
                            float a = bv;
#pragma unroll
                            for (int k = 0; k < 5; ++k) a = fmaf(wk[k], v[i + k], a);
                            o[e] = a;
                        }
                        if (p.post_elu) elu4(o[0], o[1], o[2], o[3]);
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
                    if (p.trace) tr_c += clock64() - tr3;
                    fence_proxy_async();
                    epi_bar_sync();
                    if (q == 0 && elect_one()) {
                        const CUtensorMap* mp = c == 0 ? &map_y28 : &map_y;
                        const int tc0 = c == 0 ? tcol0 + p.t_halo : tcol0 + c * 32;
                        if (p.reduce_add) tma_reduce_add_3d(mp, obuf, tc0, m_blk * BM, b);
                        else tma_store_3d(mp, obuf, tc0, m_blk * BM, b);
                        tma_commit();
                    }
                }
            }
            if (q == 0 && elect_one()) tma_wait_all();
        }
        if (p.trace && issuer) {
            unsigned long long* tr = p.trace + blockIdx.x * 24;
            tr[7] = tr_f; tr[8] = clock64() - tr_s; tr[9] = tr_d; tr[14] = tr_l; tr[15] = tr_c;
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

}  // namespace th

// ------------------------------------------------------------------------------- host side
// samples per clip a flat launch uses: T itself, or 4 for shorter rows whose pitch is 4
static inline int flat_t_of(int T, int x_rs, int y_rs) { return (T < 4 && x_rs == 4 && y_rs == 4) ? 4 : T; }

bool gemm_h_usable(const PackedMat& W, const float* X, long long x_bs, int x_rs, int T, const float* R, const float* Y,
                   long long y_bs, int y_rs, int B) {
    if (!W.H_hi || !W.H_lo) return false;
    // short chunks: flat tiles when there are enough clips (tc_flat_ok), else the flattened-column FP32 kernels.  Rows of
    // 1 - 3 samples stored with pitch 4 (every activation buffer of the codec) run as 4-sample clips: the pad columns
    // are computed and stored like any other and never read.
    if (!tc_chunk_ok(B, T) && !tc_flat_ok(B, flat_t_of(T, x_rs, y_rs))) return false;
    if ((x_rs & 3) || (x_bs & 3) || (y_rs & 3) || (y_bs & 3)) return false;
    if ((reinterpret_cast<uintptr_t>(X) & 15) || (reinterpret_cast<uintptr_t>(Y) & 15)) return false;
    if (R && (reinterpret_cast<uintptr_t>(R) & 15)) return false;
    return true;
}

// the fused DWSBlock kernel: per-clip tiles, or flat tiles at 8 samples per clip
bool gemm_h_dw_usable(const PackedMat& W, const float* X, long long x_bs, int x_rs, int T, const float* R, const float* Y,
                      long long y_bs, int y_rs, int B) {
    return gemm_h_usable(W, X, x_bs, x_rs, T, R, Y, y_bs, y_rs, B) && (tc_chunk_ok(B, T) || T == 8);
}

static int elu_poly_env() {
    static const int v = []() { const char* e = std::getenv("HILCODEC_ELU_POLY"); return (e && e[0] == '1') ? 1 : 0; }();
    return v;
}

static cudaError_t th_common(const PackedMat& W, const float* X, long long x_bs, int x_rs, int B, int T, CUtensorMap* map_hi,
                             CUtensorMap* map_lo, CUtensorMap* map_x, int* num_sms_out, bool flat = false) {
    using namespace th;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_h_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(gemm_h_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    *num_sms_out = tc::device_sm_count();
    {
        const cuuint64_t dims[2] = {(cuuint64_t)W.Kp32, (cuuint64_t)W.Mp128};
        const cuuint64_t strides[1] = {(cuuint64_t)W.Kp32 * 2};
        const cuuint32_t box[2] = {BK, BM};
        if (!tc::make_map_dt(map_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, W.H_hi, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B) ||
            !tc::make_map_dt(map_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, W.H_lo, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B))
            return cudaErrorInvalidValue;
    }
    if (flat) {   // {t, clip, k}: a box is 128 / T whole clips x 32 k; its shared-memory image is the usual [32 k][128 columns]
        const cuuint64_t dims[3] = {(cuuint64_t)T, (cuuint64_t)B, (cuuint64_t)W.K};
        const cuuint64_t strides[2] = {(cuuint64_t)x_bs * 4, (cuuint64_t)x_rs * 4};
        const cuuint32_t box[3] = {(cuuint32_t)T, (cuuint32_t)(BN / T), BK};
        if (!tc::make_map(map_x, X, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return cudaErrorInvalidValue;
    } else {
        const cuuint64_t dims[3] = {(cuuint64_t)T, (cuuint64_t)W.K, (cuuint64_t)B};
        const cuuint64_t strides[2] = {(cuuint64_t)x_rs * 4, (cuuint64_t)x_bs * 4};
        const cuuint32_t box[3] = {BN, BK, 1};
        if (!tc::make_map(map_x, X, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return cudaErrorInvalidValue;
    }
    return cudaSuccess;
}

// output map of a flat launch: {t, clip, m}, one store = 32 / T whole clips x 128 rows (a [128][32] staging chunk)
static bool flat_map_y(CUtensorMap* map_y, float* Y, long long y_bs, int y_rs, int B, int T, int M) {
    using namespace th;
    const cuuint64_t dims[3] = {(cuuint64_t)T, (cuuint64_t)B, (cuuint64_t)M};
    const cuuint64_t strides[2] = {(cuuint64_t)y_bs * 4, (cuuint64_t)y_rs * 4};
    const cuuint32_t box[3] = {(cuuint32_t)T, (cuuint32_t)(32 / T), BM};
    return tc::make_map(map_y, Y, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
}
// tile bookkeeping of a flat launch: the kernel sees ONE clip of B * T columns
static void flat_params(th::Params& p, int B, int T) {
    using namespace th;
    p.flat_t = T; p.flat_b = B;
    p.T = B * T; p.B = 1;
    p.t_step = BN; p.t_halo = 0;
    p.tiles_t = (p.T + BN - 1) / BN;
    p.total_tiles = (long long)p.num_m * p.tiles_t;
}

// weight residency (Params::a_res): on unless HILCODEC_A_RESIDENT=0, for K <= 192 and a grid that is a multiple of num_m
static unsigned long long* g_trace_buf = nullptr;
static bool trace_on() {
    static const bool on = []() { const char* e = std::getenv("HILCODEC_TRACE"); return e && e[0] == '1'; }();
    return on;
}
// HILCODEC_TRACE=1: synchronise after the launch and print the per-role wait cycles averaged over the CTAs
static void trace_report(const th::Params& p, unsigned grid, const char* what, cudaStream_t st) {
    if (!trace_on() || !g_trace_buf) return;
    cudaStreamSynchronize(st);
    std::vector<unsigned long long> h((size_t)grid * 24);
    cudaMemcpy(h.data(), g_trace_buf, h.size() * 8, cudaMemcpyDeviceToHost);
    double a[24] = {0};
    for (unsigned c = 0; c < grid; ++c)
        for (int i = 0; i < 24; ++i) a[i] += (double)h[(size_t)c * 24 + i] / grid;
    const double kbs = a[13] * ((p.K + th::BK - 1) / th::BK);
    std::fprintf(stderr,
                 "[trace] %s M=%d K=%d T=%d B=%d a_res=%d tiles/CTA=%.1f | per k-block: total %.0f | MMA waits: tempty %.0f a_full %.0f "
                 "b_ready %.0f | xform: raw_full %.0f op_empty %.0f of %.0f | epi: tfull %.0f drain %.0f of %.0f | Xprod: raw_empty %.0f of "
                 "%.0f | Aprod: a_empty %.0f | epi detail: tmem_ld %.0f taps+sts %.0f | MMA issue %.0f commit %.0f\n",
                 what, p.M, p.K, p.T, p.B, p.a_res, a[13], a[3] / kbs, a[0] / kbs, a[1] / kbs, a[2] / kbs, 2 * a[4] / kbs, 2 * a[5] / kbs,
                 2 * a[6] / kbs, a[7] / kbs, a[9] / kbs, a[8] / kbs, a[10] / kbs, a[11] / kbs, a[12] / kbs, a[14] / kbs, a[15] / kbs, a[16] / kbs, a[17] / kbs);
}

static unsigned plan_grid(th::Params& p, int num_sms) {
    using namespace th;
    if (trace_on()) {
        if (!g_trace_buf) cudaMalloc(&g_trace_buf, 1024 * 24 * sizeof(unsigned long long));
        cudaMemset(g_trace_buf, 0, 1024 * 24 * sizeof(unsigned long long));
        p.trace = g_trace_buf;
    }
    static const bool on = []() { const char* e = std::getenv("HILCODEC_A_RESIDENT"); return !(e && e[0] == '0'); }();
    const int nkb = (p.K + BK - 1) / BK;
    unsigned grid = (unsigned)(p.total_tiles < num_sms ? p.total_tiles : num_sms);
    static const bool probe = []() { const char* e = std::getenv("HILCODEC_MMA_PROBE"); return !(e && e[0] == '0'); }();
    p.probe = probe ? 1 : 0;
    p.a_res = 0;
    p.nraw = RAW_STAGES;
    p.nop = OP_STAGES;
    if (on && nkb <= A_STAGES + 2 && grid >= (unsigned)p.num_m) {
        grid = grid / p.num_m * p.num_m;
        p.a_res = 1;
        p.nop = nkb > A_STAGES ? OP_STAGES - 2 : OP_STAGES;   // k-blocks 4, 5 live in the top two operand stages
    }
    return grid;
}

static cudaError_t seed_residual(const float* R, float* Y, long long y_bs, int y_rs, int B, int M, int T, cudaStream_t st) {
    for (int b = 0; b < B; ++b) {   // out-of-place residual: seed Y with R, then accumulate in place
        cudaError_t e = cudaMemcpy2DAsync(Y + (long long)b * y_bs, (size_t)y_rs * 4, R + (long long)b * y_bs, (size_t)y_rs * 4,
                                          (size_t)T * 4, M, cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

cudaError_t launch_gemm_h(const PackedMat& W, const float* X, long long x_bs, int x_rs, int B, int T, int pre, float pre_scale,
                          const float* bias, const float* R, float* Y, long long y_bs, int y_rs, cudaStream_t st) {
    using namespace th;
    if (B == 0 || T == 0) return cudaSuccess;
    CUtensorMap map_hi, map_lo, map_x, map_y;
    int num_sms = 0;
    const bool flat = !tc_chunk_ok(B, T);
    const int Tf = flat ? flat_t_of(T, x_rs, y_rs) : T;   // samples per clip as the flat maps see them (pad columns included)
    cudaError_t e = th_common(W, X, x_bs, x_rs, B, Tf, &map_hi, &map_lo, &map_x, &num_sms, flat);
    if (e != cudaSuccess) return e;
    if (flat) {
        if (!flat_map_y(&map_y, Y, y_bs, y_rs, B, Tf, W.M)) return cudaErrorInvalidValue;
    } else {
        const cuuint64_t dims[3] = {(cuuint64_t)T, (cuuint64_t)W.M, (cuuint64_t)B};
        const cuuint64_t strides[2] = {(cuuint64_t)y_rs * 4, (cuuint64_t)y_bs * 4};
        const cuuint32_t box[3] = {32, BM, 1};
        if (!tc::make_map(&map_y, Y, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return cudaErrorInvalidValue;
    }
    if (R && R != Y) {
        e = seed_residual(R, Y, y_bs, y_rs, B, W.M, T, st);
        if (e != cudaSuccess) return e;
    }
    Params p{};
    p.M = W.M; p.K = W.K; p.T = T; p.B = B;
    p.num_m = (W.M + BM - 1) / BM;
    p.t_step = BN; p.t_halo = 0;
    p.tiles_t = (T + p.t_step - 1) / p.t_step;
    p.total_tiles = (long long)p.num_m * p.tiles_t * B;
    if (flat) flat_params(p, B, Tf);
    p.pre = pre; p.pre_scale = (pre == PRE_SCALE_ELU) ? pre_scale : 1.0f; p.bias = bias; p.reduce_add = R ? 1 : 0;
    p.c_big = W.h_inv_scale; p.c_small = W.h_inv_scale * (1.0f / LO_SCALE);
    p.xform_sleep = tc::xform_sleep_env();
    p.elu_poly = elu_poly_env();
    const unsigned grid = plan_grid(p, num_sms);
    gemm_h_kernel<false><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(map_hi, map_lo, map_x, map_y, map_y, p);
    trace_report(p, grid, "pointwise", st);
    return cudaGetLastError();
}

// Upsampling layer of the decoder in one kernel (streaming.py:633-637): act -> CausalConvTranspose1d (depthwise,
// kernel 2S, stride S, cache [B,K,1]) -> 1x1 conv + bias.  x [B][K][T_in] (low rate) -> Y [B][M][S * T_in].
bool gemm_h_up_usable(const PackedMat& W, const float* x, long long x_bs, int x_rs, int T_in, int S, int pre, const float* Y,
                      long long y_bs, int y_rs) {
    if (!W.H_hi || !W.H_lo) return false;
    if (S != 2 && S != 4 && S != 5 && S != 8) return false;
    if (pre != PRE_NONE && pre != PRE_SCALE_ELU) return false;
    if ((long long)S * T_in < 128 || (W.K & 31)) return false;
    if ((x_rs & 3) || (x_bs & 3) || (y_rs & 3) || (y_bs & 3)) return false;
    if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(Y) & 15)) return false;
    return true;
}

template <int S>
static cudaError_t launch_up(const PackedMat& W, const float* x, long long x_bs, int x_rs, int B, int T_in, int pre,
                             float pre_scale, const float* up_w, const float* ci, float* co, const float* bias, float* Y,
                             long long y_bs, int y_rs, cudaStream_t st) {
    using namespace th;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_h_kernel<false, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const int T = S * T_in;
    CUtensorMap map_hi, map_lo, map_x, map_y;
    {
        const cuuint64_t dims[2] = {(cuuint64_t)W.Kp32, (cuuint64_t)W.Mp128};
        const cuuint64_t strides[1] = {(cuuint64_t)W.Kp32 * 2};
        const cuuint32_t box[2] = {BK, BM};
        if (!tc::make_map_dt(&map_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, W.H_hi, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B) ||
            !tc::make_map_dt(&map_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, W.H_lo, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B))
            return cudaErrorInvalidValue;
    }
    {
        const cuuint64_t dims[3] = {(cuuint64_t)T_in, (cuuint64_t)W.K, (cuuint64_t)B};
        const cuuint64_t strides[2] = {(cuuint64_t)x_rs * 4, (cuuint64_t)x_bs * 4};
        const cuuint32_t box[3] = {(cuuint32_t)up_ni(S), BK, 1};
        if (!tc::make_map(&map_x, x, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return cudaErrorInvalidValue;
    }
    {
        const cuuint64_t dims[3] = {(cuuint64_t)T, (cuuint64_t)W.M, (cuuint64_t)B};
        const cuuint64_t strides[2] = {(cuuint64_t)y_rs * 4, (cuuint64_t)y_bs * 4};
        const cuuint32_t box[3] = {32, BM, 1};
        if (!tc::make_map(&map_y, Y, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return cudaErrorInvalidValue;
    }
    Params p{};
    p.M = W.M; p.K = W.K; p.T = T; p.B = B;
    p.num_m = (W.M + BM - 1) / BM;
    p.t_step = BN; p.t_halo = 0;
    p.tiles_t = (T + BN - 1) / BN;
    p.total_tiles = (long long)p.num_m * p.tiles_t * B;
    p.pre = pre; p.pre_scale = (pre == PRE_SCALE_ELU) ? pre_scale : 1.0f; p.bias = bias; p.reduce_add = 0;
    p.c_big = W.h_inv_scale; p.c_small = W.h_inv_scale * (1.0f / LO_SCALE);
    p.xform_sleep = tc::xform_sleep_env();
    p.elu_poly = 0;
    p.t_in = T_in; p.up_w = up_w; p.up_ci = ci; p.up_co = co;
    const int num_sms = tc::device_sm_count();
    const unsigned grid = plan_grid(p, num_sms);
    gemm_h_kernel<false, S><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(map_hi, map_lo, map_x, map_y, map_y, p);
    trace_report(p, grid, "upsample", st);
    return cudaGetLastError();
}

cudaError_t launch_gemm_h_up(const PackedMat& W, const float* x, long long x_bs, int x_rs, int B, int T_in, int S, int pre,
                             float pre_scale, const float* up_w, const float* cache_in, float* cache_out, const float* bias,
                             float* Y, long long y_bs, int y_rs, cudaStream_t st) {
    if (B == 0 || T_in == 0) return cudaSuccess;
#define HIL_UP(SS) \
    if (S == SS) return launch_up<SS>(W, x, x_bs, x_rs, B, T_in, pre, pre_scale, up_w, cache_in, cache_out, bias, Y, y_bs, y_rs, st);
    HIL_UP(2) HIL_UP(4) HIL_UP(5) HIL_UP(8)
#undef HIL_UP
    return cudaErrorInvalidValue;
}

// Encoder downsampling pair in one kernel: pre(x) -> 1x1 (K -> M, no bias) -> causal depthwise conv (kernel 2r, stride r) +
// bias (streaming.py:506-510).  x [B][K][T] -> Y [B][M][T / r]; cache [B][M][r] = the last r pointwise outputs.
bool gemm_h_down_usable(const PackedMat& W, const float* X, long long x_bs, int x_rs, int T, int r, const float* Y,
                        long long y_bs, int y_rs) {
    if (!W.H_hi || !W.H_lo) return false;
    if (r != 2 && r != 4 && r != 5) return false;
    if (T < 128 || (T % r) != 0) return false;
    if ((x_rs & 3) || (x_bs & 3) || (y_rs & 3) || (y_bs & 3)) return false;
    if ((reinterpret_cast<uintptr_t>(X) & 15) || (reinterpret_cast<uintptr_t>(Y) & 15)) return false;
    return true;
}

template <int R>
static cudaError_t launch_down(const PackedMat& W, const float* X, long long x_bs, int x_rs, int B, int T, int pre,
                               float pre_scale, const float* dw_w, const float* dw_b, const float* cache_in, float* cache_out,
                               float* Y, long long y_bs, int y_rs, cudaStream_t st) {
    using namespace th;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_h_kernel<false, 0, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    CUtensorMap map_hi, map_lo, map_x, map_y, map_y2;
    int num_sms = 0;
    cudaError_t e = th_common(W, X, x_bs, x_rs, B, T, &map_hi, &map_lo, &map_x, &num_sms);
    if (e != cudaSuccess) return e;
    {
        const cuuint64_t dims[3] = {(cuuint64_t)(T / R), (cuuint64_t)W.M, (cuuint64_t)B};
        const cuuint64_t strides[2] = {(cuuint64_t)y_rs * 4, (cuuint64_t)y_bs * 4};
        const cuuint32_t box1[3] = {(cuuint32_t)ds_n1(R), BM, 1};
        const cuuint32_t box2[3] = {(cuuint32_t)(ds_nout(R) - ds_n1(R)), BM, 1};
        if (!tc::make_map(&map_y, Y, 3, dims, strides, box1, CU_TENSOR_MAP_SWIZZLE_NONE) ||
            !tc::make_map(&map_y2, Y, 3, dims, strides, box2, CU_TENSOR_MAP_SWIZZLE_NONE))
            return cudaErrorInvalidValue;
    }
    Params p{};
    p.M = W.M; p.K = W.K; p.T = T; p.B = B;
    p.num_m = (W.M + BM - 1) / BM;
    p.t_step = ds_step(R); p.t_halo = ds_halo(R);
    p.tiles_t = (T + p.t_step - 1) / p.t_step;
    p.total_tiles = (long long)p.num_m * p.tiles_t * B;
    p.pre = pre; p.pre_scale = (pre == PRE_SCALE_ELU) ? pre_scale : 1.0f;
    p.dw_w = dw_w; p.dw_b = dw_b; p.cache_in = cache_in; p.cache_out = cache_out;
    p.c_big = W.h_inv_scale; p.c_small = W.h_inv_scale * (1.0f / LO_SCALE);
    p.xform_sleep = tc::xform_sleep_env();
    p.elu_poly = elu_poly_env();
    const unsigned grid = plan_grid(p, num_sms);
    gemm_h_kernel<false, 0, R><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(map_hi, map_lo, map_x, map_y, map_y2, p);
    return cudaGetLastError();
}

cudaError_t launch_gemm_h_down(const PackedMat& W, const float* X, long long x_bs, int x_rs, int B, int T, int r, int pre,
                               float pre_scale, const float* dw_w, const float* dw_b, const float* cache_in, float* cache_out,
                               float* Y, long long y_bs, int y_rs, cudaStream_t st) {
    if (B == 0 || T == 0) return cudaSuccess;
#define HIL_DOWN(RR) \
    if (r == RR) return launch_down<RR>(W, X, x_bs, x_rs, B, T, pre, pre_scale, dw_w, dw_b, cache_in, cache_out, Y, y_bs, y_rs, st);
    HIL_DOWN(2) HIL_DOWN(4) HIL_DOWN(5)
#undef HIL_DOWN
    return cudaErrorInvalidValue;
}

cudaError_t launch_gemm_h_dw(const PackedMat& W, const float* X, long long x_bs, int x_rs, int B, int T, int pre,
                             float pre_scale, const float* dw_w, const float* dw_b, const float* cache_in, float* cache_out,
                             const float* skip, float* Y, long long y_bs, int y_rs, cudaStream_t st, int post_elu) {
    if (post_elu && skip) return cudaErrorInvalidValue;   // the activation cannot follow an in-place accumulation
    using namespace th;
    if (B == 0 || T == 0) return cudaSuccess;
    CUtensorMap map_hi, map_lo, map_x, map_y, map_y28;
    int num_sms = 0;
    const bool flat = !tc_chunk_ok(B, T);
    if (flat && T != 8) return cudaErrorInvalidValue;   // the flat epilogue is written for 8 samples per clip
    cudaError_t e = th_common(W, X, x_bs, x_rs, B, T, &map_hi, &map_lo, &map_x, &num_sms, flat);
    if (e != cudaSuccess) return e;
    if (flat) {
        if (!flat_map_y(&map_y, Y, y_bs, y_rs, B, T, W.M)) return cudaErrorInvalidValue;
        map_y28 = map_y;
    } else {
        const cuuint64_t dims[3] = {(cuuint64_t)T, (cuuint64_t)W.M, (cuuint64_t)B};
        const cuuint64_t strides[2] = {(cuuint64_t)y_rs * 4, (cuuint64_t)y_bs * 4};
        const cuuint32_t box[3] = {32, BM, 1};
        const cuuint32_t box28[3] = {28, BM, 1};
        if (!tc::make_map(&map_y, Y, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B) ||
            !tc::make_map(&map_y28, Y, 3, dims, strides, box28, CU_TENSOR_MAP_SWIZZLE_NONE))
            return cudaErrorInvalidValue;
    }
    if (skip && skip != Y) {
        e = seed_residual(skip, Y, y_bs, y_rs, B, W.M, T, st);
        if (e != cudaSuccess) return e;
    }
    Params p{};
    p.M = W.M; p.K = W.K; p.T = T; p.B = B;
    p.num_m = (W.M + BM - 1) / BM;
    p.t_step = BN - 4; p.t_halo = 4;
    p.tiles_t = (T + p.t_step - 1) / p.t_step;
    p.total_tiles = (long long)p.num_m * p.tiles_t * B;
    if (flat) flat_params(p, B, T);
    p.pre = pre; p.pre_scale = (pre == PRE_SCALE_ELU) ? pre_scale : 1.0f;
    p.dw_w = dw_w; p.dw_b = dw_b; p.cache_in = cache_in; p.cache_out = cache_out;
    p.reduce_add = skip ? 1 : 0;
    p.post_elu = post_elu;
    p.c_big = W.h_inv_scale; p.c_small = W.h_inv_scale * (1.0f / LO_SCALE);
    p.xform_sleep = tc::xform_sleep_env();
    p.elu_poly = elu_poly_env();
    const unsigned grid = plan_grid(p, num_sms);
    gemm_h_kernel<true><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(map_hi, map_lo, map_x, map_y, map_y28, p);
    trace_report(p, grid, "dws", st);
    return cudaGetLastError();
}

}  // namespace hil
