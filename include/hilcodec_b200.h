/*
 * hilcodec_b200 -- C ABI of the B200-native HILCodec encode -> RVQ -> decode path.
 *
 * The reference (aask1357/hilcodec) has no FFI layer: its boundary is
 * torch.nn.Module.forward of models/hilcodec/streaming.py.  Each entry point below
 * names the reference interface it replaces (file:line relative to the reference
 * tree).  All `dev` pointers are CUDA device pointers (fp32 NCW-contiguous unless
 * stated), every call is asynchronous on `stream` (a cudaStream_t passed as void*),
 * never synchronises the host (except the *_host convenience call), returns HIL_OK or
 * a negative hil_status, and records a message retrievable with hil_last_error()
 * (thread-local).  One hil_model may be shared between threads after
 * hil_model_finalize(); a hil_state (workspace + per-stream caches for a fixed batch
 * size) must not be used from two threads at once.
 *
 * Weights are the folded deployment weights, named exactly like the reference's
 * streaming `state_dict()` keys after `remove_weight_reparameterizations()`
 * (streaming.py:740-747) with an `encoder.` / `decoder.` prefix, plus
 * `quantizer.layers.{i}.embed` -- i.e. the initializers of the published ONNX graphs.
 */
#ifndef HILCODEC_B200_H
#define HILCODEC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HIL_ABI_VERSION 1
#define HIL_MAX_STRIDES 8

typedef enum hil_status {
    HIL_OK = 0,
    HIL_ERR_INVALID = -1,   /* bad argument / shape (the reference raises from ATen or asserts) */
    HIL_ERR_MISSING = -2,   /* tensor not set before finalize */
    HIL_ERR_CUDA = -3,      /* CUDA runtime error */
    HIL_ERR_STATE = -4,     /* model not finalized / state batch mismatch */
    HIL_ERR_NOMEM = -5
} hil_status;

/* HILCodec.__init__ kwargs that shape the graph (streaming.py:651-722,
 * configs/hilcodec_{speech,music}.yaml:2-38). */
typedef struct hil_config {
    int32_t channels_enc;      /* 64  */
    int32_t channels_dec;      /* 96  */
    int32_t n_fft_base;        /* 64  */
    int32_t n_residual_enc;    /* 2   */
    int32_t n_residual_dec;    /* 3   */
    double res_scale_enc;      /* 0.5773502691896258 (python float; pre/post scales derive from it in double) */
    double res_scale_dec;      /* 0.5773502691896258 */
    int32_t n_strides;         /* 4   */
    int32_t strides[HIL_MAX_STRIDES]; /* 8,5,4,2 (decoder order; encoder uses them reversed) */
    int32_t kernel_size;       /* 5   */
    int32_t dim;               /* 128 */
    int32_t codebook_size;     /* 1024 */
    int32_t num_quantizers;    /* 8 (hil_speech) / 12 (hil_music) */
} hil_config;

typedef struct hil_model hil_model;
typedef struct hil_state hil_state;

enum { HIL_ENCODER = 0, HIL_DECODER = 1 };
/* Which of the reference's two graphs the model computes (they share all weights but differ in three places,
 * SURVEY.md section 8 quirks 1, 2, 4):
 *   HIL_GRAPH_DEPLOY  models/hilcodec/streaming.py (the ONNX export; default; what the golden vectors pin)
 *   HIL_GRAPH_TRAIN   models/hilcodec/models.py:111-118 (SEANetEncoder/Decoder, modules/seanet.py; what the
 *                     validation / PESQ loops of wrapper.py:347,373,391 call): decoder ResBlock j uses
 *                     pre_scale (1 + j*res_scale^2)^-0.5 (seanet.py:443-451), the codebook search is
 *                     argmin(-2 x.e + |e|^2) without the |x|^2 term (vector_quantize.py:132-153), and the weights
 *                     are expected with conv_post.bias * wav_std (seanet.py:464-466; fold.py graph="train"). */
enum { HIL_GRAPH_DEPLOY = 0, HIL_GRAPH_TRAIN = 1 };

int32_t hil_abi_version(void);
const char* hil_last_error(void);
void hil_config_default(hil_config* cfg, int32_t num_quantizers);

/* ---- model: replaces streaming.HILCodec(...) construction + load_state_dict ------- */
int32_t hil_model_create(const hil_config* cfg, hil_model** out);
/* host fp32 data; dims as in the reference state_dict (e.g. [C,C,1], [C,1,5], [1024,128]) */
int32_t hil_model_set_tensor(hil_model* m, const char* name, const float* host, const int64_t* dims, int32_t ndim);
/* selects the graph variant; only before hil_model_finalize (HIL_ERR_STATE afterwards) */
int32_t hil_model_set_graph(hil_model* m, int32_t graph);
int32_t hil_model_graph(const hil_model* m);
/* uploads + repacks weights for the kernels; after this the model is immutable */
int32_t hil_model_finalize(hil_model* m);
void hil_model_destroy(hil_model* m);
int32_t hil_model_hop(const hil_model* m);      /* Encoder.hop_length streaming.py:390 */
int32_t hil_model_num_caches(const hil_model* m, int32_t which); /* Encoder.num_cache streaming.py:455 (22) / 30 */
/* shape [B,C,len] of cache i in the reference's list order (streaming.py:458-470, :599-607) */
int32_t hil_model_cache_shape(const hil_model* m, int32_t which, int32_t i, int32_t batch, int64_t dims[3]);

/* ---- state: replaces HILCodec.initialize_cache (streaming.py:723-724) ------------- */
int32_t hil_state_create(hil_model* m, int32_t batch, hil_state** out);
int32_t hil_state_reset(hil_state* s, void* stream);                 /* zero all caches */
int32_t hil_state_export_cache(hil_state* s, int32_t which, int32_t i, float* dev_dst, void* stream);
int32_t hil_state_import_cache(hil_state* s, int32_t which, int32_t i, const float* dev_src, void* stream);
void hil_state_destroy(hil_state* s);
size_t hil_state_workspace_bytes(const hil_state* s);

/* ---- fp16-range guard (no counterpart in the reference, whose fp32 convolutions have no such limit) -------
 * The tensor-core kernels split every activation into two fp16 numbers, which needs |x| < 65504; beyond that the
 * split is Inf / NaN and the affected latents / PCM samples come out NaN.  The last kernel of the encoder (L2 norm)
 * and of the decoder (conv_post + tanh) set a sticky per-state device flag when they produce a non-finite value.
 * hil_state_range_flag reads it (synchronises `stream`; clear != 0 resets it).  hil_set_exact_fp32(1) makes every later
 * call of THIS thread run on the FP32 FFMA kernels (exact reference range, ~5x slower) until hil_set_exact_fp32(0);
 * returns the previous setting.  hil_state_rollback undoes the cache-generation advance of the last hil_encode /
 * hil_decode / hil_codec_forward on `s` (the previous generation is intact: caches ping-pong) so that the call can be
 * repeated.  hil_codec_forward_host does all of this by itself; the Python modules do it unless check_range is False. */
int32_t hil_state_range_flag(hil_state* s, int32_t clear, void* stream, int32_t* flag_out);
int32_t hil_state_rollback(hil_state* s, int32_t encoder, int32_t decoder);
int32_t hil_set_exact_fp32(int32_t on);

/* ---- the four calls of the deployment flow (scripts/HILCodec Onnx.ipynb cell 3) --- */
/* Encoder.forward streaming.py:482-517: wav [B,1,T] (T multiple of hop) -> z [B,T/hop,dim];
 * caches advance inside `s`. */
int32_t hil_encode(hil_model* m, hil_state* s, const float* wav_dev, int32_t B, int32_t T, float* z_dev, void* stream);
/* Same, functional list-of-tensors protocol: caches_in[22] are read, caches_out[22] written
 * (must not alias); `s` only lends its workspace. */
int32_t hil_encode_caches(hil_model* m, hil_state* s, const float* wav_dev, int32_t B, int32_t T, float* z_dev,
                          const float* const* caches_in, float* const* caches_out, void* stream);
/* SEANetEncoder.forward modules/seanet.py:368-378 for ANY length T >= 1 (one-shot, zero history): every causal
 * conv of the training graph pads itself on the right so that its last window is full (SConv1d.forward
 * modules/conv.py:222-236, get_extra_padding_for_conv1d :61-68).  wav [B,1,T] -> z [B,ceil(T/hop),dim].
 * Resets the caches in `s` first (hil_state_reset) and leaves them reset. */
int32_t hil_encode_ragged(hil_model* m, hil_state* s, const float* wav_dev, int32_t B, int32_t T, float* z_dev, void* stream);
/* ResidualVQ.forward streaming.py:89-100 (+ per-stage EuclideanCodebook.forward :51-68):
 * z [B,F,dim] -> idx [n,B,F] int64; qsum (optional) [B,F,dim] = Dequantizer(idx). */
int32_t hil_rvq_encode(hil_model* m, const float* z_dev, int32_t B, int32_t F, int32_t n, int64_t* idx_dev,
                       float* qsum_dev_or_null, void* stream);
/* Dequantizer.forward streaming.py:148-157: idx [n,B,F] int64 -> q [B,F,dim]. */
int32_t hil_rvq_decode(hil_model* m, const int64_t* idx_dev, int32_t B, int32_t F, int32_t n, float* q_dev, void* stream);
/* Decoder.forward streaming.py:619-648: q [B,F,dim] -> wav [B,1,hop*F]. */
int32_t hil_decode(hil_model* m, hil_state* s, const float* q_dev, int32_t B, int32_t F, float* wav_dev, void* stream);
int32_t hil_decode_caches(hil_model* m, hil_state* s, const float* q_dev, int32_t B, int32_t F, float* wav_dev,
                          const float* const* caches_in, float* const* caches_out, void* stream);

/* ---- fused path: what streaming.HILCodec.forward (streaming.py:726-738) means to do - */
/* wav [B,1,T] -> idx [n,B,T/hop] int64 (+ optional z [B,F,dim]) -> wav_out [B,1,T]. */
int32_t hil_codec_forward(hil_model* m, hil_state* s, const float* wav_dev, int32_t B, int32_t T, int32_t n,
                          float* z_dev_or_null, int64_t* idx_dev, float* wav_out_dev, void* stream);
/* Streaming executor: same contract as hil_codec_forward, but the launches of one step are captured into a CUDA
 * graph (keyed by the buffer pointers, T, n and the cache generation) and replayed on later calls -- a hop-sized
 * chunk is ~115 dependent launches and otherwise launch-latency bound.  Keep wav/idx/wav_out pointers fixed. */
int32_t hil_codec_forward_graph(hil_model* m, hil_state* s, const float* wav_dev, int32_t B, int32_t T, int32_t n,
                                int64_t* idx_dev, float* wav_out_dev, void* stream);
/* Same with HOST buffers (pinned recommended): H2D copy of wav, forward, D2H copy of idx and
 * wav_out, all on `stream`, then a stream synchronise.  This is the e2e call bench.py times. */
int32_t hil_codec_forward_host(hil_model* m, hil_state* s, const float* wav_host, int32_t B, int32_t T, int32_t n,
                               int64_t* idx_host, float* wav_out_host, void* stream);

/* ---- bitstream (the reference stores indices as int16 .npy, test_onnx.py:99) ------- */
/* ceil(n * log2(codebook_size) / 8): 10 bytes per frame at n = 8, 15 at n = 12 */
int32_t hil_bitstream_bytes_per_frame(const hil_model* m, int32_t n);
/* idx [n,B,F] int64 -> frame-major bytes [B*F][bytes_per_frame]; the n indices of a frame are concatenated
 * LSB first, log2(codebook_size) bits each.  Both pointers are device pointers. */
int32_t hil_pack_indices(hil_model* m, const int64_t* idx_dev, int32_t B, int32_t F, int32_t n, uint8_t* out_dev, void* stream);
int32_t hil_unpack_indices(hil_model* m, const uint8_t* in_dev, int32_t B, int32_t F, int32_t n, int64_t* idx_dev, void* stream);

/* ---- launch accounting (measurement support for bench.py; not in the reference) ----- */
/* kernels launched by this library since load (bench.py reports the per-step delta as gpu_launches) */
uint64_t hil_launch_count(void);
/* Per-kernel-category timing: between begin and end every launch is bracketed by CUDA events on
 * its stream.  Categories: 0 pointwise GEMM of the narrow layers (1x1 / fused DWSBlock / fused upsampling with
 * Cin, Cout < 384: HBM-bound), 1 STFT GEMM, 2 depthwise conv, 3 transposed depthwise, 4 conv_pre, 5 conv_post+tanh,
 * 6 RVQ, 7 misc (wav concat, l2norm, ResBlock halo gather), 8 pointwise GEMM of the wide layers (Cin or Cout >= 384:
 * bound by the tensor pipe and its operand stream), 9 fused whole-ResBlock kernel (C <= 128: HBM-bound).
 * end() synchronises the device and fills summed ms / algorithmic FLOPs / algorithmic bytes /
 * launch counts per category (arrays of HIL_PROFILE_CATEGORIES). */
/* Kernel selection for A/B measurement (returns the previous mode).  Bit 0: 1 (default) = GEMMs and STFT on the
 * tensor pipe (tcgen05, fp32-level accuracy from split operands), 0 = FP32 FFMA kernels everywhere.  Bit 2 set: do not
 * fuse DWS blocks.  Bit 4 set (default): fp16-split tensor-core kernels (gemm_h.cu, kind::f16)
 * instead of 3xTF32.  Bit 5 set: do not fuse whole ResBlocks (gemm_rb.cu).  Bit 6 set: do not fuse the decoder's
 * upsampling layers (transposed depthwise conv -> 1x1).  Bit 7 set: do not fuse the encoder's downsampling pairs
 * (1x1 -> strided depthwise conv).  Bit 8 set: residual-VQ search of batches (>= 2048 frames) on the FFMA kernel
 * instead of the tensor-core GEMM + decision kernel (bit-identical results). */
int32_t hil_set_tensor_cores(int32_t mode);
#define HIL_PROFILE_CATEGORIES 10
int32_t hil_profile_begin(void);
int32_t hil_profile_end(double* ms, double* flops, double* bytes, int64_t* launches, int32_t n_cat);
/* The launches of the last finished profile, in execution order: category, milliseconds, algorithmic FLOPs and bytes of
 * each (any pointer may be NULL; at most `cap` entries are written).  Returns the number of launches recorded -- the
 * same sequence an `ncu` launch list of the same step shows, which is how tools/summarize_launches.py attributes the
 * measured DRAM traffic to layer classes. */
int32_t hil_profile_launches(int32_t* cat, double* ms, double* flops, double* bytes, int32_t cap);

/* ---- operator level: the reference's causal primitives, for per-kernel parity tests - */
/* CausalConv1d.forward causal_layers.py:160-165, depthwise (groups=C).
 * x [B,C,T], cache_in/out [B,C,K-S], w [C,1,K], bias [C]|NULL, skip [B,C,T_out]|NULL (added),
 * pre: 0 none, 1 ELU, 2 ELU(x*pre_scale).  y [B,C,T_out], T_out=(K-S+T-K)/S+1. */
int32_t hil_op_dwconv(const float* x, const float* cache_in, float* cache_out, const float* w, const float* bias,
                      const float* skip, float* y, int32_t B, int32_t C, int32_t T, int32_t K, int32_t S,
                      int32_t pre, float pre_scale, void* stream);
/* CausalConvTranspose1d.forward causal_layers.py:183-188, depthwise K=2S, no bias.
 * x [B,C,T], cache [B,C,1], w [C,1,2S], y [B,C,S*T]. */
int32_t hil_op_dwconv_transpose(const float* x, const float* cache_in, float* cache_out, const float* w, float* y,
                                int32_t B, int32_t C, int32_t T, int32_t S, int32_t pre, float pre_scale, void* stream);
/* nn.Conv1d(k=1) built by SConv1d causal_layers.py:191-204: y[B,M,T] = W[M,K] * pre(x[B,K,T]) + bias + residual.
 * w_host is the reference-layout [M,K,1] HOST weight (packed and uploaded inside; test-only convenience). */
int32_t hil_op_pointwise(const float* x, const float* w_host, const float* bias_dev, const float* residual, float* y,
                         int32_t B, int32_t M, int32_t K, int32_t T, int32_t pre, float pre_scale, void* stream);
/* DWSBlock.forward streaming.py:189-192 (+ the ResBlock tail :268-274 when skip is given):
 * y[B,C,T] = post( dw5( W_pw[C,C] * pre(x[B,C,T]) ; cache[B,C,4] ) + b_dw + skip ).  One fused tensor-core
 * kernel for T >= 64, otherwise pointwise GEMM + depthwise kernel through tmp[B,C,T].  w_pw_host is HOST. */
int32_t hil_op_dws(const float* x, const float* w_pw_host, const float* w_dw, const float* b_dw, const float* cache_in,
                   float* cache_out, const float* skip, float* tmp, float* y, int32_t B, int32_t C, int32_t T, int32_t pre,
                   float pre_scale, int32_t post, float post_scale, void* stream);
/* ResBlock.forward streaming.py:252-275 with the residual scale folded into the second depthwise conv
 * (merge_scaling :240-250):  h[B,C,T] <- h + dw5_1(W1 * ELU(dw5_0(W0 * pre(h)) + b0)) + b1, in place.
 * c0/c1 [B,C,4] are the caches of the two depthwise convs (in -> out).  fused = 1: the one-kernel path
 * (C <= 128, C % 32 == 0, T >= 128; tmp1 receives the halo columns), fused = 0: two hil_op_dws-style launches
 * through tmp1/tmp2 [B,C,T].  w0_host / w1_host are HOST [C,C,1] weights. */
int32_t hil_op_resblock(float* h, const float* w0_host, const float* w1_host, const float* dw0_w, const float* dw0_b,
                        const float* dw1_w, const float* dw1_b, const float* c0_in, float* c0_out, const float* c1_in,
                        float* c1_out, float* tmp1, float* tmp2, int32_t B, int32_t C, int32_t T, int32_t pre, float pre_scale,
                        int32_t fused, void* stream);
/* Decoder upsampling layer, streaming.py:633-637: pre(x) -> CausalConvTranspose1d (causal_layers.py:183-188,
 * depthwise, kernel 2S, stride S, cache [B,K,1]) -> nn.Conv1d(k=1) + bias.  x [B,K,T_in] -> y [B,M,S*T_in].
 * fused = 1: one tensor-core kernel (S in {2,4,5,8}, K % 32 == 0, S*T_in >= 128, pre 0 or 2),
 * fused = 0: hil_op_dwconv_transpose + hil_op_pointwise through tmp [B,K,S*T_in] (fp32).
 * w_pw_host is a HOST [M,K,1] weight. */
int32_t hil_op_upsample(const float* x, const float* cache_in, float* cache_out, const float* w_up, const float* w_pw_host,
                        const float* bias, float* tmp, float* y, int32_t B, int32_t K, int32_t M, int32_t T_in, int32_t S,
                        int32_t pre, float pre_scale, int32_t fused, void* stream);
/* Encoder downsampling pair, streaming.py:506-510: pre(x) -> nn.Conv1d(k=1, K -> M, no bias) -> CausalConv1d
 * (causal_layers.py:160-165, depthwise, kernel 2r, stride r, cache [B,M,r]) + bias.  x [B,K,T] (T % 4 == 0 for the dense
 * rows of this entry) -> y [B,M,T/r] (T/r % 4 == 0).  fused = 1: one tensor-core kernel with the strided conv in its
 * epilogue (r in {2,4,5}, T >= 128), fused = 0: hil_op_pointwise + hil_op_dwconv through tmp [B,M,T].
 * w_pw_host is a HOST [M,K,1] weight; w_dw [M,1,2r], b_dw [M]|NULL are device pointers. */
int32_t hil_op_downsample(const float* x, const float* cache_in, float* cache_out, const float* w_pw_host, const float* w_dw,
                          const float* b_dw, float* tmp, float* y, int32_t B, int32_t K, int32_t M, int32_t T, int32_t r,
                          int32_t pre, float pre_scale, int32_t fused, void* stream);
/* CausalSTFT.forward causal_layers.py:135-144 + clamp/log streaming.py:351:
 * wav_window [B,1,(T-1)*hop+n_fft], w_host [2F,1,n_fft] HOST -> y [B,F,T] = log(max(|STFT|,1e-5)). */
int32_t hil_op_stft_logmag(const float* wav_window, const float* w_host, float* y, int32_t B, int32_t n_fft, int32_t hop,
                           int32_t T, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HILCODEC_B200_H */
