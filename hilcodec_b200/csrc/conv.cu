// Bandwidth-bound kernels of the causal conv stacks: wav-history concat, conv_pre (1->C),
// causal depthwise conv (k5 s1 and strided k=2r s=r), causal transposed depthwise conv,
// decoder conv_post (C->1) + tanh, and the encoder's L2-norm + channel-last store.
//
// Every cached conv follows causal_layers.py:160-165 / :183-188:
//     xin = cat(cache, x);  cache' = xin[..., -len(cache):];  y = conv(xin)   (no padding)
// cache_in and cache_out are distinct buffers (ping-pong), so tiles never race on them.
#include "h_split.cuh"

namespace hil {

// ------------------------------------------------------------------ wav history concat
// Encoder.forward streaming.py:486-488
__global__ void wavcat_kernel(const float* __restrict__ x, const float* __restrict__ cache_in,
                              float* __restrict__ cache_out, float* __restrict__ wav_ext, long long w_bs, int T, int P) {
    const int b = blockIdx.y;
    const int L = P + T;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < L; j += gridDim.x * blockDim.x) {
        const float v = j < P ? cache_in[(size_t)b * P + j] : x[(size_t)b * T + (j - P)];
        wav_ext[b * w_bs + j] = v;
        if (j >= T) cache_out[(size_t)b * P + (j - T)] = v;
    }
}

cudaError_t launch_wavcat(const float* x, const float* cache_in, float* cache_out, float* wav_ext, long long w_bs,
                          int B, int T, int P, cudaStream_t st) {
    if (B == 0) return cudaSuccess;
    const int L = P + T;
    dim3 grid(min((L + 255) / 256, 1024), B);
    wavcat_kernel<<<grid, 256, 0, st>>>(x, cache_in, cache_out, wav_ext, w_bs, T, P);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ conv_pre (1 -> C, k taps)
// Encoder.forward streaming.py:490 (nn.Conv1d on the last k-1 history samples + chunk)
template <int K>
__global__ void conv_pre_kernel(const float* __restrict__ win, long long w_bs, const float* __restrict__ w,
                                const float* __restrict__ bias, float* __restrict__ y, long long y_bs, int y_rs,
                                int C, int T) {
    extern __shared__ float sw[];  // [C][K] + [C]
    for (int i = threadIdx.x; i < C * K; i += blockDim.x) sw[i] = w[i];
    for (int i = threadIdx.x; i < C; i += blockDim.x) sw[C * K + i] = bias ? bias[i] : 0.f;
    __syncthreads();
    const int b = blockIdx.y;
    const int t0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (t0 >= T) return;
    float xin[K + 3];
    const float* src = win + b * w_bs + t0;
#pragma unroll
    for (int i = 0; i < K + 3; ++i) xin[i] = (t0 + i < T + K - 1) ? src[i] : 0.f;
    float* dst = y + b * y_bs + t0;
    const bool vec = (t0 + 3 < T);
    for (int c = 0; c < C; ++c) {
        float o[4];
        const float bv = sw[C * K + c];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float a = 0.f;
#pragma unroll
            for (int k = 0; k < K; ++k) a = fmaf(sw[c * K + k], xin[j + k], a);
            o[j] = a + bv;
        }
        float* d = dst + (long long)c * y_rs;
        if (vec) {
            *reinterpret_cast<float4*>(d) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
            for (int j = 0; j < 4 && t0 + j < T; ++j) d[j] = o[j];
        }
    }
}

cudaError_t launch_conv_pre(const float* win, long long w_bs, const float* w, const float* bias, float* y,
                            long long y_bs, int y_rs, int B, int C, int T, int K, cudaStream_t st) {
    if (K != 5 || (y_rs & 3) || (y_bs & 3)) return cudaErrorInvalidValue;
    if (B == 0 || T == 0) return cudaSuccess;
    dim3 grid((T + 4 * 128 - 1) / (4 * 128), B);
    conv_pre_kernel<5><<<grid, 128, (C * 5 + C) * sizeof(float), st>>>(win, w_bs, w, bias, y, y_bs, y_rs, C, T);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ causal depthwise conv
// generic: one thread per output sample.  y[t'] = bias + sum_k w[k] * xin[t'*S + k]
template <int K, int S>
__global__ void dwconv_kernel(const float* __restrict__ x, long long x_bs, int x_rs, const float* __restrict__ cache_in,
                              float* __restrict__ cache_out, const float* __restrict__ w,
                              const float* __restrict__ bias, const float* skip, float* y,
                              long long y_bs, int y_rs, int C, int T, int T_out, int pre, float pre_scale, int post,
                              float post_scale) {
    constexpr int P = K - S;
    const int c = blockIdx.y * blockDim.y + threadIdx.y, b = blockIdx.z;   // blockDim.y > 1: short chunks, many rows per CTA
    if (c >= C) return;
    const float* xr = x + b * x_bs + (long long)c * x_rs;
    const float* ci = cache_in + ((size_t)b * C + c) * P;
    float wk[K];
#pragma unroll
    for (int k = 0; k < K; ++k) wk[k] = w[c * K + k];
    const float bv = bias ? bias[c] : 0.f;
    for (int to = blockIdx.x * blockDim.x + threadIdx.x; to < T_out; to += gridDim.x * blockDim.x) {
        float a = 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int j = to * S + k;
            const float v = j < P ? ci[j] : apply_pre(xr[j - P], pre, pre_scale);
            a = fmaf(wk[k], v, a);
        }
        a += bv;
        const long long o = b * y_bs + (long long)c * y_rs + to;
        if (skip) a += skip[o];
        y[o] = apply_act_fast(a, post, post_scale);
    }
    // new cache = last P samples of xin
    if (blockIdx.x == 0 && threadIdx.x < P) {
        const int j = T + threadIdx.x;  // index into xin (length P+T)
        cache_out[((size_t)b * C + c) * P + threadIdx.x] = j < P ? ci[j] : apply_pre(xr[j - P], pre, pre_scale);
    }
}

// k=5, s=1 specialisation: 4 outputs per thread from two aligned 16-byte loads.
__global__ void dwconv5_kernel(const float* __restrict__ x, long long x_bs, int x_rs, const float* __restrict__ cache_in,
                               float* __restrict__ cache_out, const float* __restrict__ w,
                               const float* __restrict__ bias, const float* skip, float* y,
                               long long y_bs, int y_rs, int C, int T, int pre, float pre_scale, int post,
                               float post_scale) {
    const int c = blockIdx.y * blockDim.y + threadIdx.y, b = blockIdx.z;   // blockDim.y > 1: short chunks, many rows per CTA
    if (c >= C) return;
    const float* xr = x + b * x_bs + (long long)c * x_rs;
    const float* ci = cache_in + ((size_t)b * C + c) * 4;
    float wk[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) wk[k] = w[c * 5 + k];
    const float bv = bias ? bias[c] : 0.f;
    const int Tq = (T + 3) >> 2;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < Tq; q += gridDim.x * blockDim.x) {
        const int t0 = q * 4;
        float xin[8];
        if (q == 0) {
            xin[0] = ci[0]; xin[1] = ci[1]; xin[2] = ci[2]; xin[3] = ci[3];
        } else {
            const float4 v = *reinterpret_cast<const float4*>(xr + t0 - 4);
            xin[0] = apply_pre(v.x, pre, pre_scale); xin[1] = apply_pre(v.y, pre, pre_scale);
            xin[2] = apply_pre(v.z, pre, pre_scale); xin[3] = apply_pre(v.w, pre, pre_scale);
        }
        {
            // the row pitch is a multiple of 4, so this load stays inside the row's storage
            const float4 v = *reinterpret_cast<const float4*>(xr + t0);
            xin[4] = apply_pre(v.x, pre, pre_scale); xin[5] = apply_pre(v.y, pre, pre_scale);
            xin[6] = apply_pre(v.z, pre, pre_scale); xin[7] = apply_pre(v.w, pre, pre_scale);
        }
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float a = 0.f;
#pragma unroll
            for (int k = 0; k < 5; ++k) a = fmaf(wk[k], xin[j + k], a);
            o[j] = a + bv;
        }
        const long long off = b * y_bs + (long long)c * y_rs + t0;
        if (t0 + 3 < T) {
            if (skip) {
                const float4 s = *reinterpret_cast<const float4*>(skip + off);
                o[0] += s.x; o[1] += s.y; o[2] += s.z; o[3] += s.w;
            }
            if (post != PRE_NONE) {
#pragma unroll
                for (int j = 0; j < 4; ++j) o[j] = apply_act_fast(o[j], post, post_scale);
            }
            *reinterpret_cast<float4*>(y + off) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
            for (int j = 0; j < 4 && t0 + j < T; ++j)
                y[off + j] = apply_act_fast(o[j] + (skip ? skip[off + j] : 0.f), post, post_scale);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < 4) {
        const int j = T + threadIdx.x;
        cache_out[((size_t)b * C + c) * 4 + threadIdx.x] = j < 4 ? ci[j] : apply_pre(xr[j - 4], pre, pre_scale);
    }
}

// strided specialisation (encoder downsampling, k = 2r, s = r): 4 outputs per thread from aligned 16-byte loads.
// The 4 outputs to0 .. to0+3 read xin[to0*S .. to0*S + 3S + K - 1]; in x coordinates (xin = cat(cache[P], x)) that is
// x[to0*S - P .. to0*S + 4S - 1], fetched as NV float4 starting at the 16-byte boundary a0 = to0*S - ceil4(P).
template <int K, int S>
__global__ void dwconv_strided4_kernel(const float* __restrict__ x, long long x_bs, int x_rs,
                                       const float* __restrict__ cache_in, float* __restrict__ cache_out,
                                       const float* __restrict__ w, const float* __restrict__ bias, const float* skip,
                                       float* y, long long y_bs, int y_rs, int C, int T, int T_out, int pre, float pre_scale,
                                       int post, float post_scale) {
    constexpr int P = K - S;
    constexpr int P4 = (P + 3) & ~3;
    constexpr int D = P4 - P;
    constexpr int NV = (P4 + 4 * S) / 4;
    const int c = blockIdx.y * blockDim.y + threadIdx.y, b = blockIdx.z;   // blockDim.y > 1: short chunks, many rows per CTA
    if (c >= C) return;
    const float* xr = x + b * x_bs + (long long)c * x_rs;
    const float* ci = cache_in + ((size_t)b * C + c) * P;
    float wk[K];
#pragma unroll
    for (int k = 0; k < K; ++k) wk[k] = w[c * K + k];
    const float bv = bias ? bias[c] : 0.f;
    const int Tq = (T_out + 3) >> 2;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < Tq; q += gridDim.x * blockDim.x) {
        const int to0 = q * 4;
        const int a0 = to0 * S - P4;
        float buf[4 * NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int j = a0 + 4 * i;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j >= 0) {
                if (j + 3 < x_rs) {   // the row pitch is a multiple of 4: the load stays inside the row's storage
                    v = *reinterpret_cast<const float4*>(xr + j);
                    if (pre != PRE_NONE) {
                        v.x = apply_pre(v.x, pre, pre_scale); v.y = apply_pre(v.y, pre, pre_scale);
                        v.z = apply_pre(v.z, pre, pre_scale); v.w = apply_pre(v.w, pre, pre_scale);
                    }
                }
            } else {                  // history: x index j' < 0 is cache[P + j'] (only the first group of a row)
                float t[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) t[e] = (j + e >= -P) ? ci[P + j + e] : 0.f;
                v = make_float4(t[0], t[1], t[2], t[3]);
            }
            buf[4 * i] = v.x; buf[4 * i + 1] = v.y; buf[4 * i + 2] = v.z; buf[4 * i + 3] = v.w;
        }
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float a = 0.f;
#pragma unroll
            for (int k = 0; k < K; ++k) a = fmaf(wk[k], buf[D + e * S + k], a);
            o[e] = a + bv;
        }
        const long long off = b * y_bs + (long long)c * y_rs + to0;
        if (to0 + 3 < T_out) {
            if (skip) {
                const float4 sk = *reinterpret_cast<const float4*>(skip + off);
                o[0] += sk.x; o[1] += sk.y; o[2] += sk.z; o[3] += sk.w;
            }
            if (post != PRE_NONE) {
#pragma unroll
                for (int e = 0; e < 4; ++e) o[e] = apply_act_fast(o[e], post, post_scale);
            }
            *reinterpret_cast<float4*>(y + off) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
            for (int e = 0; e < 4 && to0 + e < T_out; ++e)
                y[off + e] = apply_act_fast(o[e] + (skip ? skip[off + e] : 0.f), post, post_scale);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < P) {
        const int j = T + threadIdx.x;  // index into xin (length P+T)
        cache_out[((size_t)b * C + c) * P + threadIdx.x] = j < P ? ci[j] : apply_pre(xr[j - P], pre, pre_scale);
    }
}

// Block shape of the depthwise kernels: threads along time x channel rows per CTA.  Long chunks: 32 / 128 time threads, one
// row.  Short chunks (streaming: 1 - 8 samples per row) used to launch one 32-thread CTA per (channel, clip) with one or two
// live threads -- 98 304 CTAs and 54 us for 1536 channels x 64 streams; now a CTA covers 128 / tx rows.
static inline dim3 dw_block(int time_threads, int min_tx) {
    if (time_threads >= 128) return dim3(128, 1);
    if (time_threads > 16) return dim3(32, 1);
    int tx = 1;
    while (tx < time_threads || tx < min_tx) tx <<= 1;
    return dim3(tx, 128 / tx);
}
static inline dim3 dw_grid(int time_threads, dim3 block, int C, int B) {
    return dim3(min((time_threads + (int)block.x - 1) / (int)block.x, 512), (C + (int)block.y - 1) / (int)block.y, B);
}

cudaError_t launch_dwconv(const float* x, long long x_bs, int x_rs, const float* cache_in, float* cache_out,
                          const float* w, const float* bias, const float* skip, float* y, long long y_bs, int y_rs,
                          int B, int C, int T, int K, int S, int pre, float pre_scale, int post, float post_scale,
                          cudaStream_t st) {
    if (B == 0 || C == 0) return cudaSuccess;
    if (K < S || C > 65535 || B > 65535) return cudaErrorInvalidValue;
    const int P = K - S;
    if (P + T < K) return cudaErrorInvalidValue;
    const int T_out = (P + T - K) / S + 1;
    const bool aligned = ((x_rs & 3) == 0) && ((x_bs & 3) == 0) && ((y_rs & 3) == 0) && ((y_bs & 3) == 0) &&
                         ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && ((reinterpret_cast<uintptr_t>(y) & 15) == 0) &&
                         (skip == nullptr || (reinterpret_cast<uintptr_t>(skip) & 15) == 0);
    if (K == 5 && S == 1 && aligned) {
        const int Tq = (T + 3) / 4;
        const dim3 threads = dw_block(Tq, 4), grid = dw_grid(Tq, threads, C, B);
        dwconv5_kernel<<<grid, threads, 0, st>>>(x, x_bs, x_rs, cache_in, cache_out, w, bias, skip, y, y_bs, y_rs, C, T,
                                                 pre, pre_scale, post, post_scale);
        return cudaGetLastError();
    }
    if (aligned && K == 2 * S && T % S == 0 && T_out >= 16) {
        const int Tq = (T_out + 3) / 4;
        const dim3 threads = dw_block(Tq, P), grid = dw_grid(Tq, threads, C, B);
#define HIL_DWS4(KK, SS)                                                                                                  \
    if (K == KK && S == SS) {                                                                                            \
        dwconv_strided4_kernel<KK, SS><<<grid, threads, 0, st>>>(x, x_bs, x_rs, cache_in, cache_out, w, bias, skip, y,   \
                                                                 y_bs, y_rs, C, T, T_out, pre, pre_scale, post,          \
                                                                 post_scale);                                            \
        return cudaGetLastError();                                                                                       \
    }
        HIL_DWS4(4, 2) HIL_DWS4(8, 4) HIL_DWS4(10, 5) HIL_DWS4(16, 8)
#undef HIL_DWS4
    }
    const dim3 threads = dw_block(T_out, P), grid = dw_grid(T_out, threads, C, B);
#define HIL_DW(KK, SS)                                                                                              \
    if (K == KK && S == SS) {                                                                                       \
        dwconv_kernel<KK, SS><<<grid, threads, 0, st>>>(x, x_bs, x_rs, cache_in, cache_out, w, bias, skip, y, y_bs, \
                                                         y_rs, C, T, T_out, pre, pre_scale, post, post_scale);       \
        return cudaGetLastError();                                                                                  \
    }
    HIL_DW(5, 1) HIL_DW(4, 2) HIL_DW(8, 4) HIL_DW(10, 5) HIL_DW(16, 8) HIL_DW(6, 3) HIL_DW(12, 6) HIL_DW(3, 1) HIL_DW(7, 1)
#undef HIL_DW
    return cudaErrorInvalidValue;
}

// ------------------------------------------------------------------ causal transposed depthwise
// conv_transpose1d(cat(cache[1], x), w[C,1,2S], stride S, padding S):
//   y[n] = xin[n/S + 1] * w[n%S] + xin[n/S] * w[n%S + S],  xin[0] = cache, xin[i+1] = pre(x[i])
// One thread owns 4 consecutive INPUT samples (one aligned 16-byte load + the sample before),
// applies the activation once per input (not once per output) and writes the 4*S outputs they
// generate as S aligned 16-byte stores.
template <int S>
__global__ void dwconvT_kernel(const float* __restrict__ x, long long x_bs, int x_rs, const float* __restrict__ cache_in,
                               float* __restrict__ cache_out, const float* __restrict__ w, float* __restrict__ y,
                               long long y_bs, int y_rs, int C, int T, int pre, float pre_scale, int vec) {
    const int c = blockIdx.y * blockDim.y + threadIdx.y, b = blockIdx.z;   // blockDim.y > 1: short chunks, many rows per CTA
    if (c >= C) return;
    const float* xr = x + b * x_bs + (long long)c * x_rs;
    const float cprev = cache_in[(size_t)b * C + c];
    float wc[2 * S];
#pragma unroll
    for (int k = 0; k < 2 * S; ++k) wc[k] = w[(size_t)c * 2 * S + k];
    float* yr = y + b * y_bs + (long long)c * y_rs;
    const int Tq = (T + 3) >> 2;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < Tq; q += gridDim.x * blockDim.x) {
        const int i0 = q * 4;
        float e[5];  // e[0] = xin before i0, e[1..4] = pre(x[i0..i0+3])
        e[0] = i0 == 0 ? cprev : apply_pre(xr[i0 - 1], pre, pre_scale);
        if (vec) {
            const float4 v = *reinterpret_cast<const float4*>(xr + i0);
            e[1] = apply_pre(v.x, pre, pre_scale); e[2] = apply_pre(v.y, pre, pre_scale);
            e[3] = apply_pre(v.z, pre, pre_scale); e[4] = apply_pre(v.w, pre, pre_scale);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) e[1 + j] = i0 + j < T ? apply_pre(xr[i0 + j], pre, pre_scale) : 0.f;
        }
        float o[4 * S];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int r = 0; r < S; ++r)
                // two rounded products, one rounded add (col2im-style accumulation)
                o[j * S + r] = __fadd_rn(__fmul_rn(e[1 + j], wc[r]), __fmul_rn(e[j], wc[r + S]));
        float* dst = yr + (long long)i0 * S;
        if (vec && i0 + 3 < T) {
#pragma unroll
            for (int k = 0; k < S; ++k)
                *reinterpret_cast<float4*>(dst + 4 * k) = make_float4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
        } else {
            const int nvalid = min(4, T - i0) * S;
#pragma unroll
            for (int k = 0; k < 4 * S; ++k)
                if (k < nvalid) dst[k] = o[k];
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        cache_out[(size_t)b * C + c] = T > 0 ? apply_pre(xr[T - 1], pre, pre_scale) : cprev;
}

cudaError_t launch_dwconv_transpose(const float* x, long long x_bs, int x_rs, const float* cache_in, float* cache_out,
                                    const float* w, float* y, long long y_bs, int y_rs, int B, int C, int T, int S,
                                    int pre, float pre_scale, cudaStream_t st) {
    if (B == 0 || C == 0) return cudaSuccess;
    if (C > 65535 || B > 65535 || S < 1) return cudaErrorInvalidValue;
    const int vec = ((x_rs & 3) == 0) && ((x_bs & 3) == 0) && ((y_rs & 3) == 0) && ((y_bs & 3) == 0) &&
                    ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && ((reinterpret_cast<uintptr_t>(y) & 15) == 0);
    const int Tq = (T + 3) / 4;
    const dim3 threads = dw_block(Tq, 1);
    dim3 grid = dw_grid(Tq, threads, C, B);
    if (grid.x < 1) grid.x = 1;
#define HIL_DWT(SS)                                                                                                  \
    if (S == SS) {                                                                                                   \
        dwconvT_kernel<SS><<<grid, threads, 0, st>>>(x, x_bs, x_rs, cache_in, cache_out, w, y, y_bs, y_rs, C, T, pre, \
                                                     pre_scale, vec);                                                \
        return cudaGetLastError();                                                                                   \
    }
    HIL_DWT(2) HIL_DWT(3) HIL_DWT(4) HIL_DWT(5) HIL_DWT(6) HIL_DWT(8)
#undef HIL_DWT
    return cudaErrorInvalidValue;
}

// ------------------------------------------------------------------ decoder conv_post + tanh
// Decoder.forward streaming.py:644-647: ELU -> CausalConv1d(C -> 1, k) -> Tanh
template <int K>
__global__ void conv_post_tanh_kernel(const float* __restrict__ x, long long x_bs, int x_rs,
                                      const float* __restrict__ cache_in, float* __restrict__ cache_out,
                                      const float* __restrict__ w, const float* __restrict__ bias,
                                      float* __restrict__ y, int C, int T, int pre, float pre_scale, int vec,
                                      int* __restrict__ nonfinite) {
    constexpr int P = K - 1;
    constexpr int NO = 8;          // outputs per thread: P + NO activations feed NO * K MACs per channel
    static_assert(P == 4, "the vector path loads the 4 history samples as one float4");
    extern __shared__ float sw[];  // [C][K]
    for (int i = threadIdx.x; i < C * K; i += blockDim.x) sw[i] = w[i];
    __syncthreads();
    const int b = blockIdx.y;
    const int t0 = (blockIdx.x * blockDim.x + threadIdx.x) * NO;
    const float* xb = x + b * x_bs;
    const float* cb = cache_in + (size_t)b * C * P;
    if (t0 < T) {
        float a[NO];
#pragma unroll
        for (int o = 0; o < NO; ++o) a[o] = 0.f;
        const bool full = vec && (t0 + NO <= x_rs);   // both float4 loads stay inside the row's storage
        for (int c = 0; c < C; ++c) {
            const float* xr = xb + (long long)c * x_rs;
            float xin[P + NO];  // xin index j <-> time t0 - P + j
            if (t0 == 0) {
#pragma unroll
                for (int j = 0; j < P; ++j) xin[j] = cb[c * P + j];
            } else if (vec) {
                const float4 v = *reinterpret_cast<const float4*>(xr + t0 - 4);
                xin[0] = apply_act_ex2(v.x, pre, pre_scale); xin[1] = apply_act_ex2(v.y, pre, pre_scale);
                xin[2] = apply_act_ex2(v.z, pre, pre_scale); xin[3] = apply_act_ex2(v.w, pre, pre_scale);
            } else {
#pragma unroll
                for (int j = 0; j < P; ++j) xin[j] = apply_act_ex2(xr[t0 - P + j], pre, pre_scale);
            }
            if (full) {
#pragma unroll
                for (int h = 0; h < NO / 4; ++h) {
                    const float4 v = *reinterpret_cast<const float4*>(xr + t0 + 4 * h);
                    xin[P + 4 * h + 0] = apply_act_ex2(v.x, pre, pre_scale); xin[P + 4 * h + 1] = apply_act_ex2(v.y, pre, pre_scale);
                    xin[P + 4 * h + 2] = apply_act_ex2(v.z, pre, pre_scale); xin[P + 4 * h + 3] = apply_act_ex2(v.w, pre, pre_scale);
                }
            } else {
#pragma unroll
                for (int j = 0; j < NO; ++j) xin[P + j] = t0 + j < T ? apply_act_ex2(xr[t0 + j], pre, pre_scale) : 0.f;
            }
#pragma unroll
            for (int o = 0; o < NO; ++o)
#pragma unroll
                for (int k = 0; k < K; ++k) a[o] = fmaf(sw[c * K + k], xin[o + k], a[o]);
        }
        const float bv = bias ? bias[0] : 0.f;
        bool bad = false;
        for (int o = 0; o < NO && t0 + o < T; ++o) {
            const float v = tanhf(a[o] + bv);
            bad |= !(fabsf(v) <= 1.0f);            // NaN: an activation left the fp16 range upstream (hil_state_range_flag)
            y[(size_t)b * T + t0 + o] = v;
        }
        if (bad && nonfinite) *nonfinite = 1;
    }
    if (blockIdx.x == 0) {
        for (int i = threadIdx.x; i < C * P; i += blockDim.x) {
            const int c = i / P, jj = i - c * P;
            const int j = T + jj;
            cache_out[(size_t)b * C * P + i] =
                j < P ? cb[c * P + j] : apply_act_ex2(xb[(long long)c * x_rs + (j - P)], pre, pre_scale);
        }
    }
}

// Short chunks (streaming: T = 320 per hop): the kernel above gives a thread 8 outputs and ALL C channels, which for one
// stream is 40 threads walking 96 channels one dependent load after the other (~100 us per hop, 10 % of a streaming
// frame).  Here a block is 32 time lanes x 8 channel slices: slice s sums channels s, s + 8, ... for 4 outputs per
// lane, the 8 partial sums are added in slice order (deterministic), then bias + tanh.
template <int K>
__global__ void conv_post_tanh_small_kernel(const float* __restrict__ x, long long x_bs, int x_rs,
                                            const float* __restrict__ cache_in, float* __restrict__ cache_out,
                                            const float* __restrict__ w, const float* __restrict__ bias,
                                            float* __restrict__ y, int C, int T, int pre, float pre_scale,
                                            int* __restrict__ nonfinite) {
    constexpr int P = K - 1, NO = 4, SL = 8;
    __shared__ float part[SL][32 * NO];
    const int b = blockIdx.y, tx = threadIdx.x, ty = threadIdx.y;
    const int t0 = (blockIdx.x * 32 + tx) * NO;
    const float* xb = x + b * x_bs;
    const float* cb = cache_in + (size_t)b * C * P;
    float a[NO];
#pragma unroll
    for (int o = 0; o < NO; ++o) a[o] = 0.f;
    if (t0 < T) {
        for (int c = ty; c < C; c += SL) {
            const float* xr = xb + (long long)c * x_rs;
            float xin[P + NO];   // xin index j <-> time t0 - P + j
#pragma unroll
            for (int j = 0; j < P + NO; ++j) {
                const int t = t0 - P + j;
                xin[j] = t < 0 ? cb[c * P + P + t] : (t < T ? apply_act_ex2(xr[t], pre, pre_scale) : 0.f);
            }
            float wk[K];
#pragma unroll
            for (int k = 0; k < K; ++k) wk[k] = __ldg(w + c * K + k);
#pragma unroll
            for (int o = 0; o < NO; ++o)
#pragma unroll
                for (int k = 0; k < K; ++k) a[o] = fmaf(wk[k], xin[o + k], a[o]);
        }
    }
#pragma unroll
    for (int o = 0; o < NO; ++o) part[ty][tx * NO + o] = a[o];
    __syncthreads();
    if (ty == 0 && t0 < T) {
        const float bv = bias ? bias[0] : 0.f;
        bool bad = false;
        for (int o = 0; o < NO && t0 + o < T; ++o) {
            float sum = part[0][tx * NO + o];
#pragma unroll
            for (int sl = 1; sl < SL; ++sl) sum += part[sl][tx * NO + o];
            const float v = tanhf(sum + bv);
            bad |= !(fabsf(v) <= 1.0f);
            y[(size_t)b * T + t0 + o] = v;
        }
        if (bad && nonfinite) *nonfinite = 1;
    }
    if (blockIdx.x == 0) {
        for (int i = ty * 32 + tx; i < C * P; i += 32 * SL) {
            const int c = i / P, jj = i - c * P;
            const int j = T + jj;   // index into xin = cat(cache, act(x)) of length P + T
            cache_out[(size_t)b * C * P + i] = j < P ? cb[c * P + j] : apply_act_ex2(xb[(long long)c * x_rs + (j - P)], pre, pre_scale);
        }
    }
}

cudaError_t launch_conv_post_tanh(const float* x, long long x_bs, int x_rs, const float* cache_in, float* cache_out,
                                  const float* w, const float* bias, float* y, int B, int C, int T, int K, int pre,
                                  float pre_scale, int* nonfinite, cudaStream_t st) {
    if (K != 5) return cudaErrorInvalidValue;
    if (B == 0) return cudaSuccess;
    if ((long long)B * T <= 64LL * 1024) {   // few output samples: spread the channel sum over the block
        dim3 grid((T + 127) / 128, B), block(32, 8);
        conv_post_tanh_small_kernel<5><<<grid, block, 0, st>>>(x, x_bs, x_rs, cache_in, cache_out, w, bias, y, C, T, pre,
                                                               pre_scale, nonfinite);
        return cudaGetLastError();
    }
    const int Tq = (T + 7) / 8;
    const int threads = Tq >= 128 ? 128 : 32;
    const int vec = ((x_rs & 3) == 0) && ((x_bs & 3) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    dim3 grid(max(1, (Tq + threads - 1) / threads), B);
    conv_post_tanh_kernel<5><<<grid, threads, C * 5 * sizeof(float), st>>>(x, x_bs, x_rs, cache_in, cache_out, w, bias,
                                                                           y, C, T, pre, pre_scale, vec, nonfinite);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ channel-last -> NCW
// q [B][F][C] (the dequantised latents, streaming.py:157) -> y [B][C][pitch]: lets the decoder's first 1x1 conv run on the
// tensor-core GEMM (which reads NCW boxes by TMA) for batches; 32 x 32 tiles through shared memory, both sides coalesced.
__global__ void chlast_to_ncw_kernel(const float* __restrict__ q, float* __restrict__ y, int C, int F, long long y_bs, int y_rs) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, f0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const float* qb = q + (size_t)b * F * C;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int f = f0 + r, c = c0 + threadIdx.x;
        tile[r][threadIdx.x] = (f < F && c < C) ? qb[(size_t)f * C + c] : 0.f;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int c = c0 + r, f = f0 + threadIdx.x;
        if (c < C && f < F) y[b * y_bs + (long long)c * y_rs + f] = tile[threadIdx.x][r];
    }
}

cudaError_t launch_chlast_to_ncw(const float* q, float* y, int B, int C, int F, long long y_bs, int y_rs, cudaStream_t st) {
    if (B == 0 || F == 0) return cudaSuccess;
    if (B > 65535) return cudaErrorInvalidValue;
    dim3 grid((F + 31) / 32, (C + 31) / 32, B), block(32, 8);
    chlast_to_ncw_kernel<<<grid, block, 0, st>>>(q, y, C, F, y_bs, y_rs);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ L2 norm + channel-last
// L2Norm.forward streaming.py:284-285 then x.transpose(1,2) (:517).  One warp per (b, f).
__global__ void l2norm_chlast_kernel(const float* __restrict__ x, long long x_bs, int x_rs, float* __restrict__ z,
                                     int C, int F, long long total, float scale, int* __restrict__ nonfinite) {
    const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wid >= total) return;
    const int b = (int)(wid / F), f = (int)(wid - (long long)b * F);
    const float* xb = x + b * x_bs + f;
    float ss = 0.f;
    for (int c = lane; c < C; c += 32) {
        const float v = xb[(long long)c * x_rs];
        ss = fmaf(v, v, ss);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (lane == 0 && nonfinite && !(ss <= 3.0e38f)) *nonfinite = 1;   // NaN / Inf latent (see hil_state_range_flag)
    const float denom = fmaxf(sqrtf(ss), 1e-12f);
    float* zr = z + (size_t)wid * C;
    for (int c = lane; c < C; c += 32) zr[c] = __fmul_rn(__fdiv_rn(xb[(long long)c * x_rs], denom), scale);
}

cudaError_t launch_l2norm_chlast(const float* x, long long x_bs, int x_rs, float* z, int B, int C, int F, float scale,
                                 int* nonfinite, cudaStream_t st) {
    const long long total = (long long)B * F;
    if (total == 0) return cudaSuccess;
    const long long threads = total * 32;
    l2norm_chlast_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(x, x_bs, x_rs, z, C, F, total, scale, nonfinite);
    return cudaGetLastError();
}

}  // namespace hil
