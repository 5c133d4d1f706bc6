// Skinny-N FP32 GEMM for the streaming path: the same three contracts as gemm.cu
// (launch_gemm_linear / launch_gemm_chlast_in / launch_gemm_stft_logmag), for chunks whose column count
// N = B*T is small (a hop-sized chunk of ONE stream is 1 ... 40 columns at the wide, low-rate layers).
//
// Why: gemm.cu always computes a 128 x 128 tile, so a 1536 -> 768 layer on 8 columns runs 96 dependent
// k-steps of full-tile FFMA work on 6 CTAs (~1 us each): the streaming frame was bound by kernel execution
// time, not by launch latency (CUDA-graph replay gained 4 %; profiles/r1_streaming_config4.jsonl).
//
// Mapping: one CTA = 32 output rows x NT columns (NT = 8 / 16 / 32 / 64); lane <-> row, so a warp reads
// A[k][m0 .. m0+31] of the k-major weight copy as ONE 128-byte line per k; the K range is split over the 8
// warps of the CTA; every warp stages its own activated X slice (pre() applied once per element and CTA, not
// once per lane) in shared memory in chunks of KC rows and reads it back with broadcast LDS.128; the 8 partial
// sums of an output are added in warp order 0..7 (deterministic), then bias / residual / store, or magnitude +
// clamp + log for the DFT.  Grid = (ceil(N / NT), Mp / 32): 24-48 CTAs for the wide layers instead of 6-12,
// each with K/8 sequential k per warp instead of K.
//
// Default since round 2 (measured on the B200 with the published hil_music weights, profiles/r2_ab_results.md: one
// stream 3.36 -> 1.27 ms per hop with this kernel, 0.88 ms together with the per-stage RVQ launches; 64 streams
// 4.35 -> 2.92 ms); the whole GPU parity suite runs through it wherever a chunk has <= 512 columns.
#include <cstdlib>

#include "common.cuh"

namespace hil {

namespace {

enum { SK_PLAIN = 0, SK_CHLAST = 1, SK_IM2COL = 2 };
enum { SK_LINEAR = 0, SK_LOGMAG = 1, SK_DW5 = 2 };

constexpr int SK_ROWS = 32;
constexpr int SK_WARPS = 8;

struct SkinnyParams {
    const float* A;   // [Kp][Mp] k-major
    int Mp, M, K;
    const float* X;
    long long x_bs;   // batch stride (PLAIN / IM2COL)
    long long x_ks;   // stride between consecutive k of one column
    int hop;          // IM2COL: column t starts at t*hop
    int T;
    unsigned N;       // B*T
    int pre;
    float pre_scale;
    const float* bias;
    const float* R;
    float* Y;
    long long y_bs;
    int y_rs;
    int M_out;        // rows of Y (M, or M/2 for LOGMAG)
    int cols_per_tile;  // NT, or (whole streams per tile) * T for SK_DW5
    // SK_DW5: DWSBlock tail (streaming.py:189-192 + the ResBlock add :268-274) -- causal depthwise k5 over the GEMM
    // result with its 4-sample cache, + bias, + skip, activation on the stored value (as dwconv5_kernel in conv.cu)
    const float* dw_w;      // [M][5]
    const float* dw_b;      // [M] | null
    const float* cache_in;  // [B][M][4]
    float* cache_out;       // [B][M][4]
    const float* skip;      // laid out like Y | null
    int post;
    float post_scale;
};

template <int NT, int KC, int LD, int EPI>
__global__ void __launch_bounds__(SK_WARPS * 32) skinny_kernel(const SkinnyParams p) {
    extern __shared__ __align__(16) float sk_smem[];
    float* xs_all = sk_smem;                                   // [SK_WARPS][KC][NT]
    float* red = sk_smem + SK_WARPS * KC * NT;                 // [SK_WARPS][NT][SK_ROWS + 1]
    __shared__ long long coloff[NT];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned n0 = blockIdx.x * (unsigned)p.cols_per_tile;
    const int m0 = blockIdx.y * SK_ROWS;
    const int ncol = (int)min((unsigned)p.cols_per_tile, p.N - n0);  // valid columns of this tile

    if (tid < NT) {
        long long off = 0;
        if (tid < ncol) {
            const unsigned n = n0 + tid;
            if (LD == SK_CHLAST) {
                off = (long long)n * p.K;
            } else {
                const unsigned b = n / (unsigned)p.T, t = n - b * (unsigned)p.T;
                off = (long long)b * p.x_bs + (LD == SK_IM2COL ? (long long)t * p.hop : (long long)t);
            }
        }
        coloff[tid] = off;
    }
    __syncthreads();

    // this warp's k range: chunks of KC rows dealt round-robin would interleave the partial sums; a contiguous
    // slice keeps every partial a plain sequential-k sum
    const int kper = ((p.K + SK_WARPS - 1) / SK_WARPS + KC - 1) / KC * KC;
    const int kbeg = warp * kper;
    const int kend = min(p.K, kbeg + kper);
    float* xs = xs_all + warp * KC * NT;

    float acc[NT];
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[j] = 0.f;

    // Software pipeline over chunks of KC k-rows: the weights (KC coalesced loads per lane) and the raw activations
    // (XR loads per lane) of chunk c+1 are requested before chunk c is staged and multiplied, so that a warp always
    // has one chunk of global loads in flight.
    constexpr int XR = KC * NT / 32;
    float a_cur[KC], a_nxt[KC], x_cur[XR], x_nxt[XR];
    auto load_chunk = [&](int k0, float (&a)[KC], float (&x)[XR]) {
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) {
            const int k = k0 + kk;
            a[kk] = k < kend ? __ldg(p.A + (size_t)k * p.Mp + m0 + lane) : 0.f;
        }
#pragma unroll
        for (int r = 0; r < XR; ++r) {
            const int i = lane + 32 * r;
            // PLAIN: consecutive lanes -> consecutive columns (t contiguous in memory); CHLAST / IM2COL: consecutive
            // lanes -> consecutive k (k contiguous in memory)
            const int kk = LD == SK_PLAIN ? i / NT : i % KC;
            const int j = LD == SK_PLAIN ? i % NT : i / KC;
            const int k = k0 + kk;
            float v = 0.f;
            if (k < kend && j < ncol) v = __ldg(p.X + coloff[j] + (LD == SK_PLAIN ? (long long)k * p.x_ks : (long long)k));
            x[r] = v;
        }
    };
    if (kbeg < kend) load_chunk(kbeg, a_cur, x_cur);
    for (int k0 = kbeg; k0 < kend; k0 += KC) {
        if (k0 + KC < kend) load_chunk(k0 + KC, a_nxt, x_nxt);
        // stage pre(X) of this chunk for the warp (pre() of a padded 0 may be non-zero, so padding is re-zeroed)
#pragma unroll
        for (int r = 0; r < XR; ++r) {
            const int i = lane + 32 * r;
            const int kk = LD == SK_PLAIN ? i / NT : i % KC;
            const int j = LD == SK_PLAIN ? i % NT : i / KC;
            float v = x_cur[r];
            if (LD == SK_PLAIN && p.pre != PRE_NONE) v = (k0 + kk < kend && j < ncol) ? apply_pre(v, p.pre, p.pre_scale) : 0.f;
            xs[kk * NT + j] = v;
        }
        __syncwarp();
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) {
            const float av = a_cur[kk];
#pragma unroll
            for (int q = 0; q < NT / 4; ++q) {
                const float4 x4 = *reinterpret_cast<const float4*>(&xs[kk * NT + q * 4]);
                acc[q * 4 + 0] = fmaf(av, x4.x, acc[q * 4 + 0]);
                acc[q * 4 + 1] = fmaf(av, x4.y, acc[q * 4 + 1]);
                acc[q * 4 + 2] = fmaf(av, x4.z, acc[q * 4 + 2]);
                acc[q * 4 + 3] = fmaf(av, x4.w, acc[q * 4 + 3]);
            }
        }
        __syncwarp();
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) a_cur[kk] = a_nxt[kk];
#pragma unroll
        for (int r = 0; r < XR; ++r) x_cur[r] = x_nxt[r];
    }

    // partial sums -> shared memory, fixed-order reduction over the warps
#pragma unroll
    for (int j = 0; j < NT; ++j) red[(warp * NT + j) * (SK_ROWS + 1) + lane] = acc[j];
    __syncthreads();

    if (EPI == SK_LINEAR) {
        for (int o = tid; o < SK_ROWS * ncol; o += SK_WARPS * 32) {
            const int row = o / ncol, j = o - row * ncol;
            const int m = m0 + row;
            if (m >= p.M) continue;
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < SK_WARPS; ++w) s += red[(w * NT + j) * (SK_ROWS + 1) + row];
            if (p.bias) s += p.bias[m];
            const unsigned n = n0 + j;
            const unsigned b = n / (unsigned)p.T, t = n - b * (unsigned)p.T;
            const long long yo = (long long)b * p.y_bs + (long long)m * p.y_rs + t;
            if (p.R) s += p.R[yo];
            p.Y[yo] = s;
        }
    } else if (EPI == SK_DW5) {
        // the tile holds whole streams (cols_per_tile is a multiple of T): reduce into shared memory, then slide the
        // 5-tap window along each stream's row; samples before the chunk come from the cache
        float* v = xs_all;  // [SK_ROWS][NT]; the X staging area is free after the barrier above
        for (int o = tid; o < SK_ROWS * ncol; o += SK_WARPS * 32) {
            const int row = o / ncol, j = o - row * ncol;
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < SK_WARPS; ++w) s += red[(w * NT + j) * (SK_ROWS + 1) + row];
            v[row * NT + j] = s;
        }
        __syncthreads();
        const int T = p.T;
        const unsigned b0 = n0 / (unsigned)T;
        for (int o = tid; o < SK_ROWS * ncol; o += SK_WARPS * 32) {
            const int row = o / ncol, j = o - row * ncol;
            const int m = m0 + row;
            if (m >= p.M) continue;
            const int bl = j / T, t = j - bl * T;
            const unsigned b = b0 + bl;
            const float* ci = p.cache_in + ((size_t)b * p.M + m) * 4;
            const float* vr = v + row * NT + bl * T;
            float a = 0.f;
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                const int tau = t - 4 + k;
                a = fmaf(p.dw_w[m * 5 + k], tau >= 0 ? vr[tau] : ci[4 + tau], a);
            }
            a += p.dw_b ? p.dw_b[m] : 0.f;
            const long long yo = (long long)b * p.y_bs + (long long)m * p.y_rs + t;
            if (p.skip) a += p.skip[yo];
            p.Y[yo] = apply_act_fast(a, p.post, p.post_scale);
        }
        const int nstream = ncol / T;
        for (int o = tid; o < SK_ROWS * nstream * 4; o += SK_WARPS * 32) {
            const int i = o & 3, bl = (o >> 2) % nstream, row = (o >> 2) / nstream;
            const int m = m0 + row;
            if (m >= p.M) continue;
            const unsigned b = b0 + bl;
            const int tau = T - 4 + i;  // cache_out = last 4 of cat(cache_in, v)
            p.cache_out[((size_t)b * p.M + m) * 4 + i] =
                tau >= 0 ? v[row * NT + bl * T + tau] : p.cache_in[((size_t)b * p.M + m) * 4 + 4 + tau];
        }
    } else {  // rows (2f, 2f+1) = (re_f, im_f); m0 is even
        for (int o = tid; o < (SK_ROWS / 2) * ncol; o += SK_WARPS * 32) {
            const int pr = o / ncol, j = o - pr * ncol;
            const int f = m0 / 2 + pr;
            if (f >= p.M_out) continue;
            float re = 0.f, im = 0.f;
#pragma unroll
            for (int w = 0; w < SK_WARPS; ++w) {
                re += red[(w * NT + j) * (SK_ROWS + 1) + 2 * pr];
                im += red[(w * NT + j) * (SK_ROWS + 1) + 2 * pr + 1];
            }
            // x.square().sum(dim=1).sqrt(): two rounded squares, one rounded add (as gemm.cu)
            const float mag = sqrtf(__fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im)));
            const unsigned n = n0 + j;
            const unsigned b = n / (unsigned)p.T, t = n - b * (unsigned)p.T;
            p.Y[(long long)b * p.y_bs + (long long)f * p.y_rs + t] = logf(fmaxf(mag, 1e-5f));
        }
    }
}

template <int NT, int KC, int LD, int EPI>
cudaError_t launch_one(const SkinnyParams& p, cudaStream_t st) {
    const size_t smem = (size_t)(SK_WARPS * KC * NT + SK_WARPS * NT * (SK_ROWS + 1)) * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        const cudaError_t e =
            cudaFuncSetAttribute(skinny_kernel<NT, KC, LD, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    SkinnyParams q = p;
    if (EPI == SK_DW5) q.cols_per_tile = NT / p.T * p.T;  // whole streams per tile (T <= NT checked by the caller)
    else q.cols_per_tile = NT;
    const dim3 grid((q.N + q.cols_per_tile - 1) / q.cols_per_tile, q.Mp / SK_ROWS);
    skinny_kernel<NT, KC, LD, EPI><<<grid, SK_WARPS * 32, smem, st>>>(q);
    return cudaGetLastError();
}

template <int LD, int EPI>
cudaError_t dispatch(const SkinnyParams& p, cudaStream_t st) {
    if (p.N == 0) return cudaSuccess;
    if (p.Mp % SK_ROWS) return cudaErrorInvalidValue;
    // SK_DW5 needs T <= NT; N >= T, so the choice by N guarantees it whenever T <= 64
    if (p.N <= 8) return launch_one<8, 32, LD, EPI>(p, st);
    if (p.N <= 16) return launch_one<16, 32, LD, EPI>(p, st);
    if (p.N <= 32) return launch_one<32, 16, LD, EPI>(p, st);
    return launch_one<64, 16, LD, EPI>(p, st);
}

SkinnyParams base(const PackedMat& W, int B, int T) {
    SkinnyParams p{};
    p.A = W.A; p.Mp = W.Mp; p.M = W.M; p.K = W.K; p.T = T;
    p.N = (unsigned)((long long)B * T);
    p.M_out = W.M;
    return p;
}

}  // namespace

// On by default since round 2 (one hil_music stream: 3.36 -> 0.88 ms per hop together with the per-stage RVQ launches);
// HILCODEC_SKINNY=0 falls back to gemm.cu for A/B runs, HILCODEC_SKINNY_MAXN caps the column count it is used for.
bool gemm_skinny_usable(const PackedMat& W, int B, int T) {
    static const bool on = [] { const char* e = std::getenv("HILCODEC_SKINNY"); return !(e && e[0] == '0'); }();
    static const long long max_n = [] {
        const char* e = std::getenv("HILCODEC_SKINNY_MAXN");
        return e ? std::atoll(e) : 512LL;
    }();
    const long long N = (long long)B * T;
    return on && W.A != nullptr && (W.Mp % SK_ROWS) == 0 && N > 0 && N <= max_n;
}

cudaError_t launch_gemm_skinny_linear(const PackedMat& W, const float* X, long long x_bs, int x_rs, int B, int T, int pre,
                                      float pre_scale, const float* bias, const float* R, float* Y, long long y_bs,
                                      int y_rs, cudaStream_t st) {
    SkinnyParams p = base(W, B, T);
    p.X = X; p.x_bs = x_bs; p.x_ks = x_rs; p.pre = pre; p.pre_scale = pre_scale;
    p.bias = bias; p.R = R; p.Y = Y; p.y_bs = y_bs; p.y_rs = y_rs;
    return dispatch<SK_PLAIN, SK_LINEAR>(p, st);
}

// DWSBlock for short chunks in ONE launch: y = post(dw5(W * pre(x); cache) + b_dw + skip), T <= 64
bool gemm_skinny_dws_usable(const PackedMat& W, int B, int T) { return T <= 64 && gemm_skinny_usable(W, B, T); }

cudaError_t launch_gemm_skinny_dws(const PackedMat& W, const float* X, long long x_bs, int x_rs, int B, int T, int pre,
                                   float pre_scale, const float* dw_w, const float* dw_b, const float* cache_in,
                                   float* cache_out, const float* skip, int post, float post_scale, float* Y,
                                   long long y_bs, int y_rs, cudaStream_t st) {
    if (T > 64 || !dw_w || !cache_in || !cache_out) return cudaErrorInvalidValue;
    SkinnyParams p = base(W, B, T);
    p.X = X; p.x_bs = x_bs; p.x_ks = x_rs; p.pre = pre; p.pre_scale = pre_scale;
    p.Y = Y; p.y_bs = y_bs; p.y_rs = y_rs;
    p.dw_w = dw_w; p.dw_b = dw_b; p.cache_in = cache_in; p.cache_out = cache_out; p.skip = skip;
    p.post = post; p.post_scale = post_scale;
    return dispatch<SK_PLAIN, SK_DW5>(p, st);
}

cudaError_t launch_gemm_skinny_chlast_in(const PackedMat& W, const float* Q, int B, int T, const float* bias, float* Y,
                                         long long y_bs, int y_rs, cudaStream_t st) {
    SkinnyParams p = base(W, B, T);
    p.X = Q; p.x_ks = 1; p.pre = PRE_NONE; p.bias = bias; p.Y = Y; p.y_bs = y_bs; p.y_rs = y_rs;
    return dispatch<SK_CHLAST, SK_LINEAR>(p, st);
}

cudaError_t launch_gemm_skinny_stft_logmag(const PackedMat& Wdft, const float* wav, long long w_bs, int hop, int B, int T,
                                           float* Y, long long y_bs, int y_rs, cudaStream_t st) {
    SkinnyParams p = base(Wdft, B, T);
    p.X = wav; p.x_bs = w_bs; p.x_ks = 1; p.hop = hop; p.pre = PRE_NONE;
    p.Y = Y; p.y_bs = y_bs; p.y_rs = y_rs; p.M_out = Wdft.M / 2;
    return dispatch<SK_IM2COL, SK_LOGMAG>(p, st);
}

}  // namespace hil
