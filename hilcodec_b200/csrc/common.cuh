// Shared device helpers and the internal launcher interface of libhilcodec_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdlib>

namespace hil {

enum Pre { PRE_NONE = 0, PRE_ELU = 1, PRE_SCALE_ELU = 2 };

// nn.ELU(alpha=1): x > 0 ? x : expm1(x)   (streaming.py:168-175 via activation='ELU')
__device__ __forceinline__ float elu1(float x) { return x > 0.f ? x : expm1f(x); }

__device__ __forceinline__ float apply_pre(float x, int pre, float s) {
    if (pre == PRE_NONE) return x;
    if (pre == PRE_SCALE_ELU) x = x * s;
    return elu1(x);
}

// Cheap ELU for the bandwidth-bound kernels and the GEMM transform warps (expm1f is ~30
// instructions and made them ALU-bound): for -1/16 < x <= 0 a degree-5 Taylor polynomial
// (relative error < 1e-10), below that 2^(x*log2 e) - 1 with ex2.approx (absolute error
// <= 2.4e-7).  Branch free, ~11 instructions.
__device__ __forceinline__ float elu_fast(float x) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 1.4426950408889634f));
    e -= 1.0f;
    float p = fmaf(x, 1.0f / 120.0f, 1.0f / 24.0f);
    p = fmaf(x, p, 1.0f / 6.0f);
    p = fmaf(x, p, 0.5f);
    p = fmaf(x, p, 1.0f);
    p *= x;
    const float neg = x > -0.0625f ? p : e;
    return x > 0.f ? x : neg;
}

// ELU with ex2.approx only (absolute error <= 2.4e-7, 6 instructions): the form the tensor-core transforms use
// (gemm_h.cu); also used where the activation feeds a long dot product (decoder conv_post).
__device__ __forceinline__ float elu_ex2(float x) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 1.4426950408889634f));
    return x > 0.f ? x : e - 1.0f;
}
__device__ __forceinline__ float apply_act_ex2(float x, int mode, float s) {
    if (mode == PRE_NONE) return x;
    if (mode == PRE_SCALE_ELU) x = x * s;
    return elu_ex2(x);
}

__device__ __forceinline__ float apply_act_fast(float x, int mode, float s) {
    if (mode == PRE_NONE) return x;
    if (mode == PRE_SCALE_ELU) x = x * s;
    return elu_fast(x);
}

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
static inline int pitch4(int t) { return (t + 3) & ~3; }
// Chunk lengths the per-clip 128-column tensor-core tiles are used for: T >= 32.  Measured in round 2 (hil_music, hop 320):
// giving the 40-column layers of a hop one partly filled tile per clip instead of the FP32 kernels takes one stream from
// 0.80 to 0.67 ms per hop and 64 streams from 2.9 to 2.1 ms.  HILCODEC_TC_MIN_T=<n> moves the threshold (A/B knob).
static inline bool tc_chunk_ok(int /*B*/, int T) {
    static const int min_t = []() { const char* e = std::getenv("HILCODEC_TC_MIN_T"); return e ? std::atoi(e) : 32; }();
    return T >= (min_t < 8 ? 8 : min_t);
}

// Flat tiles for SHORT chunks of MANY clips (concurrent streams: 64 streams x 8 samples): a 128-column tile is
// 128 / T whole clips, gathered by a {T, B, K} tensor map whose box is {T, 128 / T, 32} -- TMA does the flattening, the
// shared-memory image is the same [32 k][128 columns] box the kernel always sees.  Before round 2's last step these
// layers (the WIDEST of a hop: 512 - 1536 channels at 8 or 1 samples per stream) ran on the FP32 skinny kernels: 65 % of
// a 64-stream hop.  HILCODEC_TC_FLAT=0 keeps them there; fewer than `min` flat columns stay on the skinny kernels too.
static inline bool tc_flat_ok(int B, int T) {
    static const int min_cols = []() { const char* e = std::getenv("HILCODEC_TC_FLAT"); return e ? std::atoi(e) : 128; }();
    if (min_cols <= 0) return false;
    return (T == 4 || T == 8 || T == 16) && (long long)B * T >= min_cols;
}

// A weight matrix W[M][K] repacked k-major for the GEMM kernels: A[Kp][Mp], zero padded.
struct PackedMat {
    const float* A = nullptr;  // device, [Kp][Mp] for the FFMA kernels
    int M = 0, K = 0, Mp = 0, Kp = 0, TM = 8;
    // tensor-core form: row-major [Mp128][Kp32], split w = hi + lo with hi = tf32(w), lo = tf32(w - hi)
    const float* A_hi = nullptr;
    const float* A_lo = nullptr;
    int Mp128 = 0, Kp32 = 0;
    // fp16-split form for gemm_h.cu, row-major [Mp128][Kp32] halves: w * 2^s = H_hi + H_lo * 2^-11,
    // h_inv_scale = 2^-s (s chosen per matrix so that max|w * 2^s| lies in [2^13, 2^14))
    const uint16_t* H_hi = nullptr;
    const uint16_t* H_lo = nullptr;
    float h_inv_scale = 1.f;
};

// ---- gemm.cu ---------------------------------------------------------------------
int choose_tm(int M);
// Y[b][m][t] = sum_k W[m][k] * pre(X[b][k][t]) (+bias[m]) (+R[b][m][t])
cudaError_t launch_gemm_linear(const PackedMat& W, const float* X, long long x_bs, int x_rs, int B, int T,
                               int pre, float pre_scale, const float* bias, const float* R, float* Y,
                               long long y_bs, int y_rs, cudaStream_t st);
// X given channel-last: Q[(b*T+t)*K + k]  (decoder input q [B,F,dim])
cudaError_t launch_gemm_chlast_in(const PackedMat& W, const float* Q, int B, int T, const float* bias, float* Y,
                                  long long y_bs, int y_rs, cudaStream_t st);
// DFT-as-conv + magnitude + clamp + log: Wdft packed with rows interleaved (cos_f, sin_f).
// Y[b][f][t] = log(max(sqrt(re^2+im^2), 1e-5)), re/im = sum_k Wdft[.][k] * wav[b][t*hop + k]
cudaError_t launch_gemm_stft_logmag(const PackedMat& Wdft, const float* wav, long long w_bs, int hop, int B, int T,
                                    float* Y, long long y_bs, int y_rs, cudaStream_t st);

// ---- gemm_skinny.cu: the three contracts above for short chunks (streaming: <= 512 columns); HILCODEC_SKINNY=0 disables
bool gemm_skinny_usable(const PackedMat& W, int B, int T);
cudaError_t launch_gemm_skinny_linear(const PackedMat& W, const float* X, long long x_bs, int x_rs, int B, int T, int pre,
                                      float pre_scale, const float* bias, const float* R, float* Y, long long y_bs,
                                      int y_rs, cudaStream_t st);
bool gemm_skinny_dws_usable(const PackedMat& W, int B, int T);
cudaError_t launch_gemm_skinny_dws(const PackedMat& W, const float* X, long long x_bs, int x_rs, int B, int T, int pre,
                                   float pre_scale, const float* dw_w, const float* dw_b, const float* cache_in,
                                   float* cache_out, const float* skip, int post, float post_scale, float* Y,
                                   long long y_bs, int y_rs, cudaStream_t st);
cudaError_t launch_gemm_skinny_chlast_in(const PackedMat& W, const float* Q, int B, int T, const float* bias, float* Y,
                                         long long y_bs, int y_rs, cudaStream_t st);
cudaError_t launch_gemm_skinny_stft_logmag(const PackedMat& Wdft, const float* wav, long long w_bs, int hop, int B, int T,
                                           float* Y, long long y_bs, int y_rs, cudaStream_t st);

// ---- gemm_tc.cu: same contract as launch_gemm_linear, on the tensor pipe (tcgen05, 3xTF32)
bool gemm_tc_usable(const PackedMat& W, const float* X, long long x_bs, int x_rs, int T, const float* R, const float* Y,
                    long long y_bs, int y_rs, int B = 1);
cudaError_t launch_gemm_tc(const PackedMat& W, const float* X, long long x_bs, int x_rs, int B, int T, int pre,
                           float pre_scale, const float* bias, const float* R, float* Y, long long y_bs, int y_rs,
                           cudaStream_t st);

// ---- gemm_h.cu: the same two contracts with fp16 hi/lo splits (tcgen05 kind::f16, 2x the tf32 rate)
bool gemm_h_usable(const PackedMat& W, const float* X, long long x_bs, int x_rs, int T, const float* R, const float* Y,
                   long long y_bs, int y_rs, int B = 1);
bool gemm_h_dw_usable(const PackedMat& W, const float* X, long long x_bs, int x_rs, int T, const float* R, const float* Y,
                      long long y_bs, int y_rs, int B = 1);
cudaError_t launch_gemm_h(const PackedMat& W, const float* X, long long x_bs, int x_rs, int B, int T, int pre,
                          float pre_scale, const float* bias, const float* R, float* Y, long long y_bs, int y_rs,
                          cudaStream_t st);
cudaError_t launch_gemm_h_dw(const PackedMat& W, const float* X, long long x_bs, int x_rs, int B, int T, int pre,
                             float pre_scale, const float* dw_w, const float* dw_b, const float* cache_in,
                             float* cache_out, const float* skip, float* Y, long long y_bs, int y_rs, cudaStream_t st,
                             int post_elu = 0);

// encoder downsampling pair: act -> 1x1 (no bias) -> causal strided depthwise conv (kernel 2r, stride r) + bias, one kernel
bool gemm_h_down_usable(const PackedMat& W, const float* X, long long x_bs, int x_rs, int T, int r, const float* Y,
                        long long y_bs, int y_rs);
cudaError_t launch_gemm_h_down(const PackedMat& W, const float* X, long long x_bs, int x_rs, int B, int T, int r, int pre,
                               float pre_scale, const float* dw_w, const float* dw_b, const float* cache_in, float* cache_out,
                               float* Y, long long y_bs, int y_rs, cudaStream_t st);

// upsampling layer: act -> causal transposed depthwise conv (stride S, kernel 2S) -> 1x1 conv + bias, one kernel
bool gemm_h_up_usable(const PackedMat& W, const float* x, long long x_bs, int x_rs, int T_in, int S, int pre, const float* Y,
                      long long y_bs, int y_rs);
cudaError_t launch_gemm_h_up(const PackedMat& W, const float* x, long long x_bs, int x_rs, int B, int T_in, int S, int pre,
                             float pre_scale, const float* up_w, const float* cache_in, float* cache_out, const float* bias,
                             float* Y, long long y_bs, int y_rs, cudaStream_t st);

// ---- gemm_rb.cu: whole ResBlock (two DWSBlocks + residual add) in one kernel, h updated in place (C <= 128)
bool resblock_h_usable(const PackedMat& W0, const PackedMat& W1, const float* h, long long bs, int rs, int T);
size_t resblock_h_halo_floats(int C, int T, int B);
cudaError_t launch_resblock_halo(const float* h, long long bs, int rs, int B, int C, int T, float* halo, cudaStream_t st);
cudaError_t launch_resblock_h(const PackedMat& W0, const PackedMat& W1, float* h, long long bs, int rs, int B, int T, int pre,
                              float pre_scale, const float* dw0_w, const float* dw0_b, const float* dw1_w, const float* dw1_b,
                              const float* c0_in, float* c0_out, const float* c1_in, float* c1_out, const float* halo,
                              cudaStream_t st);

// fused DWSBlock: y = dw5(W * pre(x)) + b_dw (+ skip); same usability conditions as launch_gemm_tc
cudaError_t launch_gemm_tc_dw(const PackedMat& W, const float* X, long long x_bs, int x_rs, int B, int T, int pre,
                              float pre_scale, const float* dw_w, const float* dw_b, const float* cache_in,
                              float* cache_out, const float* skip, float* Y, long long y_bs, int y_rs, cudaStream_t st);

// ---- stft_tc.cu: launch_gemm_stft_logmag on the tensor pipe
bool stft_tc_usable(const PackedMat& Wdft, const float* wav, long long w_bs, int T, const float* Y, long long y_bs,
                    int y_rs, int B = 1, int hop = 1);
cudaError_t launch_stft_tc(const PackedMat& Wdft, const float* wav, long long w_bs, int hop, int B, int T, float* Y,
                           long long y_bs, int y_rs, cudaStream_t st);

// ---- conv.cu ---------------------------------------------------------------------
// wav_ext[b][0:P+T] = cat(cache_in[b][0:P], x[b][0:T]); cache_out = last P of it.
cudaError_t launch_wavcat(const float* x, const float* cache_in, float* cache_out, float* wav_ext, long long w_bs,
                          int B, int T, int P, cudaStream_t st);
// conv_pre: y[b][co][t] = bias[co] + sum_k w[co][k] * win[b][t+k]   (1 -> C, dense k taps)
cudaError_t launch_conv_pre(const float* win, long long w_bs, const float* w, const float* bias, float* y,
                            long long y_bs, int y_rs, int B, int C, int T, int K, cudaStream_t st);
// causal depthwise conv, see hil_op_dwconv; post: activation applied to the stored output (after bias + skip)
cudaError_t launch_dwconv(const float* x, long long x_bs, int x_rs, const float* cache_in, float* cache_out,
                          const float* w, const float* bias, const float* skip, float* y, long long y_bs, int y_rs,
                          int B, int C, int T, int K, int S, int pre, float pre_scale, int post, float post_scale,
                          cudaStream_t st);
cudaError_t launch_dwconv_transpose(const float* x, long long x_bs, int x_rs, const float* cache_in, float* cache_out,
                                    const float* w, float* y, long long y_bs, int y_rs, int B, int C, int T, int S,
                                    int pre, float pre_scale, cudaStream_t st);
// decoder conv_post: y[b][t] = tanh(bias + sum_c sum_k w[c][k] * xin[b][c][t+k]), xin = cat(cache, pre(x))
cudaError_t launch_conv_post_tanh(const float* x, long long x_bs, int x_rs, const float* cache_in, float* cache_out,
                                  const float* w, const float* bias, float* y, int B, int C, int T, int K, int pre,
                                  float pre_scale, int* nonfinite, cudaStream_t st);
// q [B][F][C] -> y [B][C][pitch]
cudaError_t launch_chlast_to_ncw(const float* q, float* y, int B, int C, int F, long long y_bs, int y_rs, cudaStream_t st);
// z[b][f][c] = x[b][c][f] / max(||x[b][:,f]||, 1e-12) * scale
// nonfinite (may be null): set to 1 when a latent / PCM sample comes out NaN or Inf
cudaError_t launch_l2norm_chlast(const float* x, long long x_bs, int x_rs, float* z, int B, int C, int F, float scale,
                                 int* nonfinite, cudaStream_t st);

// ---- rvq.cu ----------------------------------------------------------------------
// codebooks [n_q][size][dim], ee [n_q][size] = sum_k e^2
cudaError_t launch_codebook_norms(const float* codebooks, float* ee, int n_q, int size, int dim, cudaStream_t st);
cudaError_t launch_rvq_encode(const float* z, const float* codebooks, const float* ee, int size, int dim, long long frames,
                              int n, int64_t* idx, float* qsum, bool drop_xx, cudaStream_t st);
// few-frame (streaming) variant: one launch per stage over (code tiles) x (frame blocks); HILCODEC_RVQ_SPLIT=0 disables
bool rvq_split_usable(int size, int dim, long long frames);
size_t rvq_split_scratch_bytes();
// tensor-core batch search (rvq.cu): per-stage decision kernel on the GEMM's dot products + the final transpose
cudaError_t launch_rvq_tc_select(const float* Y, float* Rk, float* Qk, const float* cb, const float* ees, int size,
                                 long long pitch, long long frames, int first, int64_t* idx, float ee_max, bool drop_xx,
                                 unsigned int* rescored, cudaStream_t st);
cudaError_t launch_kmajor_to_rows(const float* x, long long pitch, float* y, int C, long long F, cudaStream_t st);
cudaError_t launch_rvq_encode_split(const float* z, const float* codebooks, const float* ee, int size, int dim,
                                    long long frames, int n, int64_t* idx, float* qsum, bool drop_xx, void* scratch,
                                    cudaStream_t st);
// few-frame variant in ONE launch: a cluster of (code tiles) CTAs per 32 frames, candidates exchanged through distributed
// shared memory; HILCODEC_RVQ_CLUSTER=0 disables
bool rvq_cluster_usable(int size, int dim, long long frames);
cudaError_t launch_rvq_encode_cluster(const float* z, const float* codebooks, const float* ee, int size, int dim,
                                      long long frames, int n, int64_t* idx, float* qsum, bool drop_xx, cudaStream_t st);
cudaError_t launch_rvq_decode(const int64_t* idx, const float* codebooks, int size, int dim, long long frames, int n,
                              float* q, cudaStream_t st);

// ---- bitpack.cu: indices [n][frames] int64 <-> frame-major bitstream, `bits` per index, LSB first
cudaError_t launch_pack_indices(const int64_t* idx, long long frames, int n, int bits, uint8_t* out, cudaStream_t st);
cudaError_t launch_unpack_indices(const uint8_t* in, long long frames, int n, int bits, int64_t* idx, cudaStream_t st);

}  // namespace hil
