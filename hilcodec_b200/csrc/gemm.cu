// FP32 (FFMA) tiled GEMM for the 1x1 "pointwise" convolutions and the DFT-as-conv STFT.
//
//   Y[b][m][t] = epi( sum_k A[k][m] * pro(X[b][k][t]) )
//
// The (b,t) plane is flattened to one column index n = b*T + t, so low-rate layers
// (T = 75 per clip) still fill 128-wide column tiles.  A is the weight matrix repacked
// k-major at model-finalize time ([Kp][Mp], zero padded), so both operands stream into
// shared memory with 16-byte loads and the inner product runs from registers:
// 256 threads, 16x16 thread grid, TM x 8 accumulators per thread, BK = 16, double buffered.
//
// Replaces: nn.Conv1d(k=1) from SConv1d (causal_layers.py:191-204), the ELU in front of it
// (streaming.py:168-175), the bias / residual add after it (streaming.py:365), and
// CausalSTFT + clamp + log (causal_layers.py:135-144, streaming.py:346-351).
#include "common.cuh"

namespace hil {

enum { LD_PLAIN = 0, LD_CHLAST = 1, LD_IM2COL = 2 };
enum { EPI_LINEAR = 0, EPI_LOGMAG = 1 };

struct GemmParams {
    const float* A;
    int Mp, M, K;
    const float* X;
    long long x_bs;   // batch stride (PLAIN / IM2COL)
    int x_rs;         // row (k) stride (PLAIN)
    int hop;          // IM2COL
    int T;
    unsigned N;       // B*T
    int pre;
    float pre_scale;
    const float* bias;
    const float* R;
    float* Y;
    long long y_bs;
    int y_rs;
    int M_out;        // rows of Y (M, or F for LOGMAG)
};

constexpr int BN = 128;
constexpr int BK = 16;

template <int TM, int LD, int EPI>
__global__ void __launch_bounds__(256, 2) gemm_kernel(const GemmParams p) {
    constexpr int BM = 16 * TM;
    __shared__ __align__(16) float As[2][BK][BM];
    __shared__ __align__(16) float Bs[2][BK][BN];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const unsigned n0 = blockIdx.x * (unsigned)BN;
    const int m0 = blockIdx.y * BM;
    const int T = p.T;

    // ---- B-operand loader state (fixed columns per thread for the whole K loop)
    long long boff[4];
    bool bvalid[4];
    bool bvec = false;
    int b_kr;  // first k row this thread loads
    if (LD == LD_CHLAST) {
        const unsigned n = n0 + (tid & 127);
        bvalid[0] = n < p.N;
        boff[0] = (long long)n * p.K;
        b_kr = tid >> 7;
    } else {
        const int cg = tid & 31;
        b_kr = tid >> 5;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const unsigned n = n0 + cg * 4 + j;
            bvalid[j] = n < p.N;
            const unsigned b = bvalid[j] ? n / (unsigned)T : 0u;
            const unsigned t = bvalid[j] ? n - b * (unsigned)T : 0u;
            boff[j] = (LD == LD_IM2COL) ? (long long)b * p.x_bs + (long long)t * p.hop
                                         : (long long)b * p.x_bs + t;
        }
        if (LD == LD_PLAIN) {
            bvec = bvalid[3] && (boff[3] - boff[0] == 3) && ((p.x_rs & 3) == 0) &&
                   ((reinterpret_cast<uintptr_t>(p.X + boff[0]) & 15) == 0);
        }
    }

    float4 areg[2];
    float4 breg[2];
    constexpr int A_VEC = BK * BM / 4;  // float4 per A tile

    auto load_tile = [&](int kt) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int i = tid + r * 256;
            if (i < A_VEC) {
                const int kk = i / (BM / 4), m4 = i - kk * (BM / 4);
                areg[r] = *reinterpret_cast<const float4*>(p.A + (size_t)(kt * BK + kk) * p.Mp + m0 + m4 * 4);
            }
        }
        if (LD == LD_CHLAST) {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int k = kt * BK + (b_kr + 2 * r) * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (bvalid[0] && k < p.K) v = *reinterpret_cast<const float4*>(p.X + boff[0] + k);
                breg[r] = v;
            }
        } else {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int k = kt * BK + b_kr + 8 * r;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (k < p.K) {
                    if (LD == LD_IM2COL) {
                        if (bvalid[0]) v.x = p.X[boff[0] + k];
                        if (bvalid[1]) v.y = p.X[boff[1] + k];
                        if (bvalid[2]) v.z = p.X[boff[2] + k];
                        if (bvalid[3]) v.w = p.X[boff[3] + k];
                    } else {
                        const long long ko = (long long)k * p.x_rs;
                        if (bvec) {
                            v = *reinterpret_cast<const float4*>(p.X + boff[0] + ko);
                        } else {
                            if (bvalid[0]) v.x = p.X[boff[0] + ko];
                            if (bvalid[1]) v.y = p.X[boff[1] + ko];
                            if (bvalid[2]) v.z = p.X[boff[2] + ko];
                            if (bvalid[3]) v.w = p.X[boff[3] + ko];
                        }
                        if (p.pre != PRE_NONE) {
                            v.x = apply_pre(v.x, p.pre, p.pre_scale);
                            v.y = apply_pre(v.y, p.pre, p.pre_scale);
                            v.z = apply_pre(v.z, p.pre, p.pre_scale);
                            v.w = apply_pre(v.w, p.pre, p.pre_scale);
                        }
                    }
                }
                breg[r] = v;
            }
        }
    };

    auto store_tile = [&](int buf) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int i = tid + r * 256;
            if (i < A_VEC) {
                const int kk = i / (BM / 4), m4 = i - kk * (BM / 4);
                *reinterpret_cast<float4*>(&As[buf][kk][m4 * 4]) = areg[r];
            }
        }
        if (LD == LD_CHLAST) {
            const int col = tid & 127;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int kk = (b_kr + 2 * r) * 4;
                Bs[buf][kk + 0][col] = breg[r].x;
                Bs[buf][kk + 1][col] = breg[r].y;
                Bs[buf][kk + 2][col] = breg[r].z;
                Bs[buf][kk + 3][col] = breg[r].w;
            }
        } else {
            const int cg = tid & 31;
#pragma unroll
            for (int r = 0; r < 2; ++r)
                *reinterpret_cast<float4*>(&Bs[buf][b_kr + 8 * r][cg * 4]) = breg[r];
        }
    };

    float acc[TM][8];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    const int nk = (p.K + BK - 1) / BK;
    load_tile(0);
    store_tile(0);
    __syncthreads();
    int cur = 0;
    for (int kt = 0; kt < nk; ++kt) {
        const bool more = kt + 1 < nk;
        if (more) load_tile(kt + 1);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[TM], b[8];
            if (TM % 4 == 0) {
#pragma unroll
                for (int q = 0; q < TM / 4; ++q) {
                    const float4 v = *reinterpret_cast<const float4*>(&As[cur][kk][ty * TM + q * 4]);
                    a[q * 4 + 0] = v.x; a[q * 4 + 1] = v.y; a[q * 4 + 2] = v.z; a[q * 4 + 3] = v.w;
                }
            } else {
#pragma unroll
                for (int q = 0; q < TM / 2; ++q) {
                    const float2 v = *reinterpret_cast<const float2*>(&As[cur][kk][ty * TM + q * 2]);
                    a[q * 2 + 0] = v.x; a[q * 2 + 1] = v.y;
                }
            }
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][kk][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[cur][kk][64 + tx * 4]);
            b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
            b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (more) store_tile(cur ^ 1);
        __syncthreads();
        cur ^= 1;
    }

    // ---- epilogue
#pragma unroll
    for (int g = 0; g < 2; ++g) {
        long long yoff[4];
        bool yvalid[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const unsigned n = n0 + g * 64 + tx * 4 + j;
            yvalid[j] = n < p.N;
            const unsigned b = yvalid[j] ? n / (unsigned)T : 0u;
            const unsigned t = yvalid[j] ? n - b * (unsigned)T : 0u;
            yoff[j] = (long long)b * p.y_bs + t;
        }
        const bool yvec = yvalid[3] && (yoff[3] - yoff[0] == 3) && ((p.y_rs & 3) == 0) &&
                          ((reinterpret_cast<uintptr_t>(p.Y + yoff[0]) & 15) == 0) &&
                          (p.R == nullptr || (reinterpret_cast<uintptr_t>(p.R + yoff[0]) & 15) == 0);
        if (EPI == EPI_LINEAR) {
#pragma unroll
            for (int i = 0; i < TM; ++i) {
                const int m = m0 + ty * TM + i;
                if (m >= p.M) continue;
                const float bv = p.bias ? p.bias[m] : 0.f;
                const long long ro = (long long)m * p.y_rs;
                float v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = acc[i][g * 4 + j] + bv;
                if (yvec) {
                    if (p.R) {
                        const float4 r = *reinterpret_cast<const float4*>(p.R + yoff[0] + ro);
                        v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w;
                    }
                    *reinterpret_cast<float4*>(p.Y + yoff[0] + ro) = make_float4(v[0], v[1], v[2], v[3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (!yvalid[j]) continue;
                        float o = v[j];
                        if (p.R) o += p.R[yoff[j] + ro];
                        p.Y[yoff[j] + ro] = o;
                    }
                }
            }
        } else {  // EPI_LOGMAG: rows (2f, 2f+1) = (re_f, im_f)
#pragma unroll
            for (int i = 0; i < TM / 2; ++i) {
                const int f = (m0 + ty * TM) / 2 + i;
                if (f >= p.M_out) continue;
                const long long ro = (long long)f * p.y_rs;
                float v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float re = acc[2 * i][g * 4 + j], im = acc[2 * i + 1][g * 4 + j];
                    // x.square().sum(dim=1).sqrt(): two rounded squares, one rounded add
                    const float mag = sqrtf(__fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im)));
                    v[j] = logf(fmaxf(mag, 1e-5f));
                }
                if (yvec) {
                    *reinterpret_cast<float4*>(p.Y + yoff[0] + ro) = make_float4(v[0], v[1], v[2], v[3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (yvalid[j]) p.Y[yoff[j] + ro] = v[j];
                }
            }
        }
    }
}

int choose_tm(int M) {
    if (M % 128 == 0) return 8;
    if (M % 96 == 0) return 6;
    if (M % 64 == 0) return 4;
    if (M > 256) return 8;
    return M > 96 ? 8 : (M > 64 ? 6 : 4);
}

template <int LD, int EPI>
static cudaError_t dispatch(const PackedMat& W, const GemmParams& p, cudaStream_t st) {
    if (p.N == 0) return cudaSuccess;
    const int BM = 16 * W.TM;
    dim3 grid((p.N + BN - 1) / BN, W.Mp / BM);
    switch (W.TM) {
        case 8: gemm_kernel<8, LD, EPI><<<grid, 256, 0, st>>>(p); break;
        case 6: gemm_kernel<6, LD, EPI><<<grid, 256, 0, st>>>(p); break;
        case 4: gemm_kernel<4, LD, EPI><<<grid, 256, 0, st>>>(p); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

static GemmParams base_params(const PackedMat& W, int B, int T) {
    GemmParams p{};
    p.A = W.A; p.Mp = W.Mp; p.M = W.M; p.K = W.K; p.T = T;
    p.N = (unsigned)((long long)B * T);
    p.M_out = W.M;
    return p;
}

cudaError_t launch_gemm_linear(const PackedMat& W, const float* X, long long x_bs, int x_rs, int B, int T, int pre,
                               float pre_scale, const float* bias, const float* R, float* Y, long long y_bs, int y_rs,
                               cudaStream_t st) {
    if ((long long)B * T >= (1LL << 31)) return cudaErrorInvalidValue;
    GemmParams p = base_params(W, B, T);
    p.X = X; p.x_bs = x_bs; p.x_rs = x_rs; p.pre = pre; p.pre_scale = pre_scale;
    p.bias = bias; p.R = R; p.Y = Y; p.y_bs = y_bs; p.y_rs = y_rs;
    return dispatch<LD_PLAIN, EPI_LINEAR>(W, p, st);
}

cudaError_t launch_gemm_chlast_in(const PackedMat& W, const float* Q, int B, int T, const float* bias, float* Y,
                                  long long y_bs, int y_rs, cudaStream_t st) {
    if ((long long)B * T >= (1LL << 31) || (W.K & 3)) return cudaErrorInvalidValue;
    GemmParams p = base_params(W, B, T);
    p.X = Q; p.pre = PRE_NONE; p.bias = bias; p.Y = Y; p.y_bs = y_bs; p.y_rs = y_rs;
    return dispatch<LD_CHLAST, EPI_LINEAR>(W, p, st);
}

cudaError_t launch_gemm_stft_logmag(const PackedMat& Wdft, const float* wav, long long w_bs, int hop, int B, int T,
                                    float* Y, long long y_bs, int y_rs, cudaStream_t st) {
    if ((long long)B * T >= (1LL << 31) || Wdft.TM != 6) return cudaErrorInvalidValue;
    GemmParams p = base_params(Wdft, B, T);
    p.X = wav; p.x_bs = w_bs; p.hop = hop; p.pre = PRE_NONE;
    p.Y = Y; p.y_bs = y_bs; p.y_rs = y_rs; p.M_out = Wdft.M / 2;
    if (p.N == 0) return cudaSuccess;
    dim3 grid((p.N + BN - 1) / BN, Wdft.Mp / 96);
    gemm_kernel<6, LD_IM2COL, EPI_LOGMAG><<<grid, 256, 0, st>>>(p);
    return cudaGetLastError();
}

}  // namespace hil
