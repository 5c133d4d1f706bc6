// Residual vector quantizer: L2 codebook search, residual subtract and in-order dequant sum
// for all n stages in ONE kernel (the residual never leaves the SM).
//
// Replaces ResidualVQ.forward (streaming.py:89-100) with EuclideanCodebook.forward
// (streaming.py:51-68) per stage, and Dequantizer.forward (streaming.py:148-157).
//
//   dist[c] = -((sum_k r_k^2 - 2 * sum_k r_k e_ck) + sum_k e_ck^2);  idx = argmax (first wins)
//   r <- r - E[idx];   qsum <- qsum + E[idx]   (stage order, fp32)
// drop_xx = the training graph's search (models/hilcodec/vector_quantize.py:146-152):
//   distance[c] = (-2 r) . e_c + sum_k e_ck^2;  idx = argmin (first wins)
// which is the same expression with sum_k r_k^2 replaced by 0 (0 - a = -a and the outer negation are exact).
//
// Layout: warp w of a CTA owns FPW consecutive frames (so the per-frame argmax reduction and the
// residual update are warp-local shuffles); lane l scores codes l, l+32, l+64, l+96 of each
// 128-code tile staged in shared memory.
#include <cstdlib>

#include "common.cuh"

namespace hil {

constexpr int RVQ_SPLIT_MAX_FRAMES = 1024;
constexpr int RVQ_SPLIT_MAX_TILES = 8;
constexpr int RVQ_FT = 32;      // frames per CTA
constexpr int RVQ_CT = 128;     // codes per smem tile
constexpr int RVQ_DIM = 128;    // vector dimension (fixed by the kernel's lane mapping)
constexpr int RVQ_PITCH = RVQ_DIM + 4;

__global__ void codebook_norm_kernel(const float* __restrict__ cb, float* __restrict__ ee, long long rows, int dim) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const float* e = cb + r * dim;
    float s = 0.f;
    for (int k = 0; k < dim; ++k) s = __fadd_rn(s, __fmul_rn(e[k], e[k]));
    ee[r] = s;
}

cudaError_t launch_codebook_norms(const float* codebooks, float* ee, int n_q, int size, int dim, cudaStream_t st) {
    const long long rows = (long long)n_q * size;
    if (rows == 0) return cudaSuccess;
    codebook_norm_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, st>>>(codebooks, ee, rows, dim);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------
// The one-kernel search (all n stages, residual in shared memory).  Round 1's kernel gave a warp 4 frames and a CTA 32;
// this one (measured in round 2: 2.92 -> 2.22 ms at 19 200 frames x 12 stages, bit-identical indices and sums) differs
// in two ways, same arithmetic per frame:
//  * FPW = 8 frames per warp: per k4 step a warp issues 8 broadcast LDS.128 (residuals) + 4 four-wavefront
//    LDS.128 (codes) = 24 shared-memory wavefronts for 128 FFMA instead of 20 for 64, which moves the inner loop from
//    the LDS pipe (1 wavefront / clk / SM) to the FMA pipes (4 warp-FFMA / clk / SM);
//  * the number of warps per CTA is chosen by the launcher so that ONE wave of 2 CTAs per SM covers all frames
//    (config 3: 19 200 frames = 2400 warp units over 296 CTA slots -> 9 warps per CTA, 267 CTAs) instead of 600
//    32-frame CTAs running as 2.03 -> 3 rounds.
template <int FPW>
__global__ void __launch_bounds__(288, 2)
rvq_encode_kernel(const float* __restrict__ z, const float* __restrict__ codebooks, const float* __restrict__ ee,
                     int size, long long frames, int n, int64_t* __restrict__ idx, float* __restrict__ qsum, int drop_xx) {
    extern __shared__ __align__(16) float smem[];
    const int nthreads = blockDim.x, nwarps = nthreads >> 5;
    const int FT = nwarps * FPW;                   // frames per CTA
    float* R = smem;                               // [FT][RVQ_PITCH]
    float* E = smem + (size_t)FT * RVQ_PITCH;      // [RVQ_CT][RVQ_PITCH]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long f0 = (long long)blockIdx.x * FT;

    for (int i = tid; i < FT * (RVQ_DIM / 4); i += nthreads) {
        const int fr = i / (RVQ_DIM / 4), k4 = i - fr * (RVQ_DIM / 4);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (f0 + fr < frames) v = *reinterpret_cast<const float4*>(z + (f0 + fr) * RVQ_DIM + k4 * 4);
        *reinterpret_cast<float4*>(&R[fr * RVQ_PITCH + k4 * 4]) = v;
    }
    __syncthreads();

    for (int s = 0; s < n; ++s) {
        const float* cb = codebooks + (size_t)s * size * RVQ_DIM;
        const float* ees = ee + (size_t)s * size;

        float xx[FPW];
#pragma unroll
        for (int f = 0; f < FPW; ++f) {
            const float4 v = *reinterpret_cast<const float4*>(&R[(warp * FPW + f) * RVQ_PITCH + lane * 4]);
            float p = __fmul_rn(v.x, v.x);
            p = fmaf(v.y, v.y, p); p = fmaf(v.z, v.z, p); p = fmaf(v.w, v.w, p);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
            xx[f] = drop_xx ? 0.f : p;
        }

        float best[FPW];
        int besti[FPW];
#pragma unroll
        for (int f = 0; f < FPW; ++f) { best[f] = -INFINITY; besti[f] = 0x7fffffff; }

        for (int c0 = 0; c0 < size; c0 += RVQ_CT) {
            __syncthreads();  // previous tile fully consumed
            for (int i = tid; i < RVQ_CT * (RVQ_DIM / 4); i += nthreads) {
                const int cr = i / (RVQ_DIM / 4), k4 = i - cr * (RVQ_DIM / 4);
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (c0 + cr < size) v = *reinterpret_cast<const float4*>(cb + (size_t)(c0 + cr) * RVQ_DIM + k4 * 4);
                *reinterpret_cast<float4*>(&E[cr * RVQ_PITCH + k4 * 4]) = v;
            }
            __syncthreads();

            float dot[FPW][4];
#pragma unroll
            for (int f = 0; f < FPW; ++f)
#pragma unroll
                for (int j = 0; j < 4; ++j) dot[f][j] = 0.f;
#pragma unroll 2
            for (int k4 = 0; k4 < RVQ_DIM / 4; ++k4) {
                float4 e4[4];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    e4[j] = *reinterpret_cast<const float4*>(&E[(lane + 32 * j) * RVQ_PITCH + k4 * 4]);
#pragma unroll
                for (int f = 0; f < FPW; ++f) {
                    const float4 r4 = *reinterpret_cast<const float4*>(&R[(warp * FPW + f) * RVQ_PITCH + k4 * 4]);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float d = dot[f][j];   // the same sequential-k FMA chain as rvq_encode_kernel
                        d = fmaf(r4.x, e4[j].x, d);
                        d = fmaf(r4.y, e4[j].y, d);
                        d = fmaf(r4.z, e4[j].z, d);
                        d = fmaf(r4.w, e4[j].w, d);
                        dot[f][j] = d;
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int code = c0 + lane + 32 * j;
                if (code < size) {
                    const float e2 = ees[code];
#pragma unroll
                    for (int f = 0; f < FPW; ++f) {
                        const float d = -__fadd_rn(__fsub_rn(xx[f], __fmul_rn(2.f, dot[f][j])), e2);
                        if (d > best[f]) { best[f] = d; besti[f] = code; }
                    }
                }
            }
        }

#pragma unroll
        for (int f = 0; f < FPW; ++f) {
            float bd = best[f];
            int bi = besti[f];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float od = __shfl_xor_sync(0xffffffffu, bd, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (od > bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
            }
            if (bi < 0 || bi >= size) bi = 0;
            const long long fr = f0 + warp * FPW + f;
            const float4 e = *reinterpret_cast<const float4*>(cb + (size_t)bi * RVQ_DIM + lane * 4);
            float4* rp = reinterpret_cast<float4*>(&R[(warp * FPW + f) * RVQ_PITCH + lane * 4]);
            float4 r = *rp;
            r.x = __fsub_rn(r.x, e.x); r.y = __fsub_rn(r.y, e.y); r.z = __fsub_rn(r.z, e.z); r.w = __fsub_rn(r.w, e.w);
            *rp = r;
            if (qsum && fr < frames) {
                // the dequantised sum lives in its output row between stages (this lane's 16 bytes, L2-resident) instead
                // of in 32 registers: q = ((0 + e_0) + e_1) + ... in stage order, as the one-kernel search
                float4* qp = reinterpret_cast<float4*>(qsum + fr * RVQ_DIM + lane * 4);
                float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
                if (s > 0) q = *qp;
                q.x = __fadd_rn(q.x, e.x); q.y = __fadd_rn(q.y, e.y); q.z = __fadd_rn(q.z, e.z); q.w = __fadd_rn(q.w, e.w);
                *qp = q;
            }
            if (lane == 0 && fr < frames) idx[(size_t)s * frames + fr] = bi;
        }
        __syncwarp();
    }

}

// warps per CTA such that one wave of `slots` CTAs covers all warp units when possible (<= 9 warps = 288 threads keeps
// two CTAs per SM within the register and shared-memory budgets), otherwise 8
int rvq_v2_warps(long long frames, int fpw, int slots) {
    const long long units = (frames + fpw - 1) / fpw;
    const long long w = (units + slots - 1) / slots;
    if (w <= 9) return (int)(w < 1 ? 1 : w);
    return 8;
}

static cudaError_t launch_rvq_encode_big(const float* z, const float* codebooks, const float* ee, int size, long long frames,
                                        int n, int64_t* idx, float* qsum, bool drop_xx, cudaStream_t st) {
    constexpr int FPW = 8;
    static int slots = 0;
    if (!slots) {
        int dev = 0, sms = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        slots = 2 * (sms > 0 ? sms : 148);
        const size_t max_smem = (size_t)(9 * FPW + RVQ_CT) * RVQ_PITCH * sizeof(float);
        const cudaError_t e = cudaFuncSetAttribute(rvq_encode_kernel<FPW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   (int)max_smem);
        if (e != cudaSuccess) { slots = 0; return e; }
    }
    const int warps = rvq_v2_warps(frames, FPW, slots);
    const int ft = warps * FPW;
    const size_t smem = (size_t)(ft + RVQ_CT) * RVQ_PITCH * sizeof(float);
    const unsigned grid = (unsigned)((frames + ft - 1) / ft);
    rvq_encode_kernel<FPW><<<grid, warps * 32, smem, st>>>(z, codebooks, ee, size, frames, n, idx, qsum, drop_xx ? 1 : 0);
    return cudaGetLastError();
}

cudaError_t launch_rvq_encode(const float* z, const float* codebooks, const float* ee, int size, int dim, long long frames,
                              int n, int64_t* idx, float* qsum, bool drop_xx, cudaStream_t st) {
    if (dim != RVQ_DIM) return cudaErrorInvalidValue;
    if (frames == 0 || n == 0) return cudaSuccess;
    return launch_rvq_encode_big(z, codebooks, ee, size, frames, n, idx, qsum, drop_xx, st);
}

// ------------------------------------------------------------------------------------------------------------
// Tensor-core search for BATCHES (round 2; codec.cu run_rvq_tc).  Per stage the 2 * frames * size * 128 FLOP of dot
// products are ONE fp32-accurate tensor-core GEMM (gemm_h.cu, the same kernel as the 1x1 convolutions):
//     Y[f / 128][c][f % 128] = sum_k E_s[c][k] * Rk[k][f]      Rk = the residuals, k-major [128][pitch]; the GEMM is
//                                                                launched with one "clip" per 128-frame tile, so the
//                                                                dot products of a tile are one contiguous 512 KB block
// and the kernel below turns Y into the stage's indices with the EXACT arithmetic of rvq_encode_kernel wherever the
// tensor-core value could change the decision: v[c] = 2 Y[c][f] - ee[c] is scanned for the best and
// second-best code; when they are closer than tau (a bound on |v - exact| + the FFMA chain's own rounding, 20x the
// observed error) every code within tau of the best is re-scored with the sequential-k fp32 FMA chain and the
// reference's -((xx - 2 dot) + ee) expression, first index winning ties -- so indices, residuals and sums are
// bit-identical to the one-kernel search (tests/test_gpu_ops.py, test_kernel_emulation.py).  The residual update and
// the running dequantised sum are the same fp32 subtract / add per element, in stage order, on the k-major copies.
constexpr int RVQ_TC_TILE = 128;   // frames per Y block = the GEMM's column tile
constexpr int RVQ_TC_CPW = 256;    // codes per warp of the decision kernel: codebooks of <= 2048 codes

__global__ void __launch_bounds__(256, 5)
rvq_tc_select_kernel(const float* __restrict__ Y, float* __restrict__ Rk, float* __restrict__ Qk,
                     const float* __restrict__ cb, const float* __restrict__ ees, int size, long long pitch,
                     long long frames, int first, int64_t* __restrict__ idx, float ee_max, int drop_xx,
                     unsigned int* __restrict__ rescored) {
    // CTA = 32 frames (lane = frame, so every row access is one 128-byte line) x 8 warps; warp w owns residual dims
    // 16w .. 16w+15 and scans codes [w * size/8, (w+1) * size/8)
    __shared__ float sm_p[32][33];
    __shared__ float sm_ee[8][RVQ_TC_CPW];
    __shared__ float sm_xx[32], sm_lim[32];
    __shared__ float sm_v1[8][32], sm_v2[8][32];
    __shared__ int sm_i1[8][32];
    __shared__ int sm_idx[32], sm_flag[32];
    __shared__ int sm_nflag;
    __shared__ float sm_wd[8];
    __shared__ int sm_wi[8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long f0 = (long long)blockIdx.x * 32;
    const bool live = f0 + lane < frames;
    const long long f = live ? f0 + lane : frames - 1;   // idle lanes read a valid column and write nothing
    // dot products of frame f, tile-major: Y[f / 128][code][f % 128] (one 512 KB block per 128 frames)
    const float* y = Y + (f / RVQ_TC_TILE) * (long long)size * RVQ_TC_TILE + (f % RVQ_TC_TILE);

    // |r|^2 with rvq_encode_kernel's association: partials over 4 dims, then the xor-butterfly tree
    float rv[16];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int i = 0; i < 4; ++i) rv[4 * j + i] = Rk[(long long)(16 * w + 4 * j + i) * pitch + f];
        float q = __fmul_rn(rv[4 * j], rv[4 * j]);
        q = fmaf(rv[4 * j + 1], rv[4 * j + 1], q);
        q = fmaf(rv[4 * j + 2], rv[4 * j + 2], q);
        q = fmaf(rv[4 * j + 3], rv[4 * j + 3], q);
        sm_p[4 * w + j][lane] = q;
    }
    if (threadIdx.x == 0) sm_nflag = 0;
    __syncthreads();
    if (w == 0) {
        float p[32];
#pragma unroll
        for (int l = 0; l < 32; ++l) p[l] = sm_p[l][lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int l = 0; l < o; ++l) p[l] = __fadd_rn(p[l], p[l + o]);
        sm_xx[lane] = p[0];
    }

    {   // best and second-best of v[c] = 2 Y[c][f] - ee[c] over this warp's codes (ascending: the first maximum wins);
        // 8 independent loads are issued before their compare chain; the warp's |e|^2 values sit in shared memory
        const int cpw = (size + 7) / 8;
        const int c_lo = w * cpw, c_hi = min(size, c_lo + cpw);
        for (int i = lane; i < c_hi - c_lo; i += 32) sm_ee[w][i] = __ldg(ees + c_lo + i);
        __syncwarp();
        float v1 = -INFINITY, v2 = -INFINITY;
        int i1 = 0x7fffffff;
        // (branch-free second best: v2 = max(v2, min(v, v1)) before v1 moves; 32-bit offsets with immediate strides --
        // the first version spent 29 instructions per code and lane, ncu: 17.7 M warp instructions per launch)
        const float* yp = y + (long long)c_lo * RVQ_TC_TILE;
        const float* ep = sm_ee[w];
        const int ncode = c_hi - c_lo;
        int c = 0;
        for (; c + 8 <= ncode; c += 8) {
            float yv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) yv[u] = yp[(c + u) * RVQ_TC_TILE];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float v = fmaf(2.f, yv[u], -ep[c + u]);
                v2 = fmaxf(v2, fminf(v, v1));
                if (v > v1) { v1 = v; i1 = c_lo + c + u; }
            }
        }
        for (; c < ncode; ++c) {
            const float v = fmaf(2.f, yp[c * RVQ_TC_TILE], -ep[c]);
            v2 = fmaxf(v2, fminf(v, v1));
            if (v > v1) { v1 = v; i1 = c_lo + c; }
        }
        sm_v1[w][lane] = v1; sm_v2[w][lane] = v2; sm_i1[w][lane] = i1;
    }
    __syncthreads();
    if (w == 0) {
        float v1 = -INFINITY, v2 = -INFINITY;
        int i1 = 0x7fffffff;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const float a1 = sm_v1[u][lane], a2 = sm_v2[u][lane];
            if (a1 > v1) { v2 = fmaxf(v1, a2); v1 = a1; i1 = sm_i1[u][lane]; }
            else v2 = fmaxf(v2, a1);
        }
        const float tau = 4e-5f * (sm_xx[lane] + ee_max);
        if (live && !(v1 - v2 >= tau)) {   // near tie (or NaN): decided below with the exact expression
            const int slot = atomicAdd(&sm_nflag, 1);
            sm_flag[slot] = lane;
            sm_lim[lane] = v1 - tau;
        }
        sm_idx[lane] = i1;
    }
    __syncthreads();

    const int nflag = sm_nflag;
    for (int q = 0; q < nflag; ++q) {
        const int fl = sm_flag[q];
        const float lim = sm_lim[fl];
        const float xs = drop_xx ? 0.f : sm_xx[fl];
        const float* yq = Y + ((f0 + fl) / RVQ_TC_TILE) * (long long)size * RVQ_TC_TILE + ((f0 + fl) % RVQ_TC_TILE);
        const float* r = Rk + f0 + fl;
        float bd = -INFINITY;
        int bi = 0x7fffffff;
        for (int c = threadIdx.x; c < size; c += 256) {
            const float v = fmaf(2.f, yq[(long long)c * RVQ_TC_TILE], -__ldg(ees + c));
            if (v < lim) continue;
            const float4* e = reinterpret_cast<const float4*>(cb + (size_t)c * RVQ_DIM);
            float dot = 0.f;   // the sequential-k FMA chain of rvq_encode_kernel
            for (int k4 = 0; k4 < RVQ_DIM / 4; ++k4) {
                const float4 e4 = __ldg(e + k4);
                dot = fmaf(r[(long long)(4 * k4) * pitch], e4.x, dot);
                dot = fmaf(r[(long long)(4 * k4 + 1) * pitch], e4.y, dot);
                dot = fmaf(r[(long long)(4 * k4 + 2) * pitch], e4.z, dot);
                dot = fmaf(r[(long long)(4 * k4 + 3) * pitch], e4.w, dot);
            }
            const float d = -__fadd_rn(__fsub_rn(xs, __fmul_rn(2.f, dot)), __ldg(ees + c));
            if (d > bd) { bd = d; bi = c; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float od = __shfl_xor_sync(0xffffffffu, bd, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (od > bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
        }
        if (lane == 0) { sm_wd[w] = bd; sm_wi[w] = bi; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int u = 1; u < 8; ++u)
                if (sm_wd[u] > bd || (sm_wd[u] == bd && sm_wi[u] < bi)) { bd = sm_wd[u]; bi = sm_wi[u]; }
            sm_idx[fl] = bi;
            if (rescored) atomicAdd(rescored, 1u);
        }
        __syncthreads();
    }

    int i1 = sm_idx[lane];
    if (i1 < 0 || i1 >= size) i1 = 0;
    if (live) {
        const float4* e = reinterpret_cast<const float4*>(cb + (size_t)i1 * RVQ_DIM + 16 * w);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float4 e4 = __ldg(e + j);
            const float ev[4] = {e4.x, e4.y, e4.z, e4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const long long o = (long long)(16 * w + 4 * j + i) * pitch + f;
                Rk[o] = __fsub_rn(rv[4 * j + i], ev[i]);
                if (Qk) Qk[o] = __fadd_rn(first ? 0.f : Qk[o], ev[i]);
            }
        }
        if (w == 0) idx[f] = i1;
    }
}

cudaError_t launch_rvq_tc_select(const float* Y, float* Rk, float* Qk, const float* cb, const float* ees, int size,
                                 long long pitch, long long frames, int first, int64_t* idx, float ee_max, bool drop_xx,
                                 unsigned int* rescored, cudaStream_t st) {
    if (frames == 0) return cudaSuccess;
    rvq_tc_select_kernel<<<(unsigned)((frames + 31) / 32), 256, 0, st>>>(Y, Rk, Qk, cb, ees, size, pitch, frames, first,
                                                                         idx, ee_max, drop_xx ? 1 : 0, rescored);
    return cudaGetLastError();
}

// k-major [C][pitch] -> rows [F][C] (the dequantised sum back in the layout the decoder and the callers read)
__global__ void kmajor_to_rows_kernel(const float* __restrict__ x, long long pitch, float* __restrict__ y, int C, long long F) {
    __shared__ float tile[32][33];
    const long long f0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    for (int rr = threadIdx.y; rr < 32; rr += blockDim.y) {
        const int c = c0 + rr;
        const long long f = f0 + threadIdx.x;
        tile[rr][threadIdx.x] = (c < C && f < F) ? x[(long long)c * pitch + f] : 0.f;
    }
    __syncthreads();
    for (int rr = threadIdx.y; rr < 32; rr += blockDim.y) {
        const long long f = f0 + rr;
        const int c = c0 + threadIdx.x;
        if (f < F && c < C) y[f * C + c] = tile[threadIdx.x][rr];
    }
}

cudaError_t launch_kmajor_to_rows(const float* x, long long pitch, float* y, int C, long long F, cudaStream_t st) {
    if (F == 0) return cudaSuccess;
    dim3 grid((unsigned)((F + 31) / 32), (C + 31) / 32), block(32, 8);
    kmajor_to_rows_kernel<<<grid, block, 0, st>>>(x, pitch, y, C, F);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------
// Split variant for FEW frames (streaming: one frame per stream and hop).  The kernel above gives 32 frames to one
// CTA, which then walks all n stages x 8 code tiles alone (~3.5 us per tile => ~340 us for n = 12 however few frames
// there are).  Here stage s is ONE launch of (code tiles) x (frame blocks) CTAs: every CTA first finishes stage s-1
// (picks the best of the per-tile candidates the previous launch left in `part_in`, lowest code index on ties, and
// rebuilds the residual r = ((z - E_0[i_0]) - E_1[i_1]) ... in stage order, i.e. the same fp32 operations as above),
// then scores ITS 128 codes of stage s with the same FMA order as above and leaves (best distance, index) per frame
// in `part_out`.  The residual entering stage s and the running dequantised sum are handed from launch to launch through
// two ping-pong scratch rows per frame (launch s reads side (s-1) & 1, its tile-0 CTA writes side s & 1), so every
// launch does ONE gather (round 2: rebuilding both from z with s gathers per launch made the chain O(n^2) dependent
// L2 round trips, 180 us of a 880 us streaming hop).  A last launch (s == n) writes the final index and the sum.  n + 1 launches of ~4 us
// instead of one of ~340 us; results are bit-identical to rvq_encode_kernel.  Default for <= 1024 frames since round 2
// (one hil_music stream: 1.27 -> 0.88 ms per hop; HILCODEC_RVQ_SPLIT=0 keeps the one-kernel search for A/B runs).
struct RvqCand { float d; int i; };

__global__ void __launch_bounds__(256, 2)
rvq_stage_kernel(const float* __restrict__ z, const float* __restrict__ codebooks, const float* __restrict__ ee, int size,
                 int tiles, long long frames, int s, int n, int64_t* __restrict__ idx, float* __restrict__ qsum,
                 const RvqCand* __restrict__ part_in, RvqCand* __restrict__ part_out, float* __restrict__ rq_scratch,
                 int drop_xx) {
    extern __shared__ __align__(16) float smem[];
    float* R = smem;                              // [RVQ_FT][RVQ_PITCH]
    float* E = smem + RVQ_FT * RVQ_PITCH;         // [RVQ_CT][RVQ_PITCH]
    // rq_scratch: [2 sides][residual | qsum][RVQ_SPLIT_MAX_FRAMES][RVQ_DIM]
    const size_t side = (size_t)2 * RVQ_SPLIT_MAX_FRAMES * RVQ_DIM;
    const float* r_in = rq_scratch + (size_t)((s - 1) & 1) * side;
    const float* q_in = r_in + (size_t)RVQ_SPLIT_MAX_FRAMES * RVQ_DIM;
    float* r_out = rq_scratch + (size_t)(s & 1) * side;
    float* q_out = r_out + (size_t)RVQ_SPLIT_MAX_FRAMES * RVQ_DIM;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long f0 = (long long)blockIdx.y * RVQ_FT;
    const int tile = blockIdx.x;
    const bool last = s == n;

    // ---- finish stage s-1 and rebuild the residual of this warp's 4 frames (lane owns dims 4*lane .. 4*lane+3)
#pragma unroll
    for (int f = 0; f < 4; ++f) {
        const long long fr = f0 + warp * 4 + f;
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        if (fr < frames) {
            // residual / sum entering stage s-1 (z / 0 for the first stage), then stage s-1's code: the same fp32
            // operations in the same order as the one-kernel search
            const bool first = s <= 1;
            r = *reinterpret_cast<const float4*>((first ? z : r_in) + fr * RVQ_DIM + lane * 4);
            if (!first) q = *reinterpret_cast<const float4*>(q_in + fr * RVQ_DIM + lane * 4);
            if (s > 0) {
                const int j = s - 1;
                float bd = -INFINITY;
                int bi = 0x7fffffff;
                for (int t = 0; t < tiles; ++t) {       // ascending code ranges: strict > keeps the first maximum
                    const RvqCand c = part_in[(size_t)fr * tiles + t];
                    if (c.d > bd || (c.d == bd && c.i < bi)) { bd = c.d; bi = c.i; }
                }
                if (bi < 0 || bi >= size) bi = 0;       // all-NaN row, as above
                if (tile == 0 && lane == 0) idx[(size_t)j * frames + fr] = bi;
                const float4 e = *reinterpret_cast<const float4*>(codebooks + ((size_t)j * size + bi) * RVQ_DIM + lane * 4);
                r.x = __fsub_rn(r.x, e.x); r.y = __fsub_rn(r.y, e.y); r.z = __fsub_rn(r.z, e.z); r.w = __fsub_rn(r.w, e.w);
                q.x = __fadd_rn(q.x, e.x); q.y = __fadd_rn(q.y, e.y); q.z = __fadd_rn(q.z, e.z); q.w = __fadd_rn(q.w, e.w);
                if (tile == 0 && !last) {
                    *reinterpret_cast<float4*>(r_out + fr * RVQ_DIM + lane * 4) = r;
                    *reinterpret_cast<float4*>(q_out + fr * RVQ_DIM + lane * 4) = q;
                }
            }
            if (last && tile == 0 && qsum) *reinterpret_cast<float4*>(qsum + fr * RVQ_DIM + lane * 4) = q;
        }
        *reinterpret_cast<float4*>(&R[(warp * 4 + f) * RVQ_PITCH + lane * 4]) = r;
    }
    if (last) return;

    const float* cb = codebooks + (size_t)s * size * RVQ_DIM;
    const float* ees = ee + (size_t)s * size;
    const int c0 = tile * RVQ_CT;

    // this CTA's code tile
    for (int i = tid; i < RVQ_CT * (RVQ_DIM / 4); i += 256) {
        const int cr = i / (RVQ_DIM / 4), k4 = i - cr * (RVQ_DIM / 4);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c0 + cr < size) v = *reinterpret_cast<const float4*>(cb + (size_t)(c0 + cr) * RVQ_DIM + k4 * 4);
        *reinterpret_cast<float4*>(&E[cr * RVQ_PITCH + k4 * 4]) = v;
    }
    __syncthreads();  // R rows of this warp and the E tile are complete

    // xx[f], identical to rvq_encode_kernel
    float xx[4];
#pragma unroll
    for (int f = 0; f < 4; ++f) {
        const float4 v = *reinterpret_cast<const float4*>(&R[(warp * 4 + f) * RVQ_PITCH + lane * 4]);
        float p = __fmul_rn(v.x, v.x);
        p = fmaf(v.y, v.y, p); p = fmaf(v.z, v.z, p); p = fmaf(v.w, v.w, p);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
        xx[f] = drop_xx ? 0.f : p;
    }

    float dot[4][4];
#pragma unroll
    for (int f = 0; f < 4; ++f)
#pragma unroll
        for (int j = 0; j < 4; ++j) dot[f][j] = 0.f;
#pragma unroll 4
    for (int k4 = 0; k4 < RVQ_DIM / 4; ++k4) {
        float4 r4[4], e4[4];
#pragma unroll
        for (int f = 0; f < 4; ++f) r4[f] = *reinterpret_cast<const float4*>(&R[(warp * 4 + f) * RVQ_PITCH + k4 * 4]);
#pragma unroll
        for (int j = 0; j < 4; ++j) e4[j] = *reinterpret_cast<const float4*>(&E[(lane + 32 * j) * RVQ_PITCH + k4 * 4]);
#pragma unroll
        for (int f = 0; f < 4; ++f)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float d = dot[f][j];
                d = fmaf(r4[f].x, e4[j].x, d);
                d = fmaf(r4[f].y, e4[j].y, d);
                d = fmaf(r4[f].z, e4[j].z, d);
                d = fmaf(r4[f].w, e4[j].w, d);
                dot[f][j] = d;
            }
    }
    float best[4];
    int besti[4];
#pragma unroll
    for (int f = 0; f < 4; ++f) { best[f] = -INFINITY; besti[f] = 0x7fffffff; }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int code = c0 + lane + 32 * j;
        if (code < size) {
            const float e2 = ees[code];
#pragma unroll
            for (int f = 0; f < 4; ++f) {
                const float d = -__fadd_rn(__fsub_rn(xx[f], __fmul_rn(2.f, dot[f][j])), e2);
                if (d > best[f]) { best[f] = d; besti[f] = code; }
            }
        }
    }
#pragma unroll
    for (int f = 0; f < 4; ++f) {
        float bd = best[f];
        int bi = besti[f];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float od = __shfl_xor_sync(0xffffffffu, bd, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (od > bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
        }
        const long long fr = f0 + warp * 4 + f;
        if (lane == 0 && fr < frames) {
            RvqCand c;
            c.d = bd; c.i = bi;
            part_out[(size_t)fr * tiles + tile] = c;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// Cluster variant of the few-frame search (streaming): the n + 1 per-stage launches above cost ~12 us each on a
// one-stream hop (160 us of a 670 us hop).  Here ONE launch runs all n stages: a thread-block cluster of `tiles`
// CTAs (8 for 1024 codes; the portable maximum) owns 32 frames, CTA r scores code tile r exactly like
// rvq_stage_kernel, sends its (best distance, index) per frame to EVERY CTA of the cluster through distributed shared
// memory, and after one cluster barrier each CTA picks the winner (same rule: largest distance, lowest index on ties),
// gathers that code and updates its own copy of the residual.  The next stage's code tile is prefetched with cp.async
// while the current one is scored.  Same fp32 operations in the same order as the one-kernel search: bit-identical.
__device__ __forceinline__ uint32_t rvq_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void rvq_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void rvq_cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

// FPW = frames per warp (1, 2 or 4): a CTA covers 8 * FPW frames.  Few frames (64 concurrent streams = 64 frames per hop)
// are spread over more clusters with FPW = 1 instead of leaving 3 of 4 frame slots per warp to two clusters.
template <int FPW>
__global__ void __launch_bounds__(256, 1)
rvq_cluster_kernel(const float* __restrict__ z, const float* __restrict__ codebooks, const float* __restrict__ ee, int size,
                   int tiles, long long frames, int n, int64_t* __restrict__ idx, float* __restrict__ qsum, int drop_xx) {
    constexpr int FT = 8 * FPW;                             // frames per CTA
    extern __shared__ __align__(16) float smem[];
    float* R = smem;                                        // [FT][RVQ_PITCH]
    float* E0 = smem + FT * RVQ_PITCH;                  // 2 x [RVQ_CT][RVQ_PITCH]
    RvqCand* cand = reinterpret_cast<RvqCand*>(E0 + 2 * RVQ_CT * RVQ_PITCH);   // [2 parities][tiles][FT]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long f0 = (long long)blockIdx.y * FT;
    const int tile = blockIdx.x;                            // = rank in the cluster (cluster dims (tiles, 1, 1))
    const int c0 = tile * RVQ_CT;

    auto prefetch = [&](int s) {                            // this CTA's code tile of stage s -> E[s & 1]
        const float* cb = codebooks + (size_t)s * size * RVQ_DIM;
        float* E = E0 + (size_t)(s & 1) * RVQ_CT * RVQ_PITCH;
        for (int i = tid; i < RVQ_CT * (RVQ_DIM / 4); i += 256) {
            const int cr = i / (RVQ_DIM / 4), k4 = i - cr * (RVQ_DIM / 4);
            float* dst = &E[cr * RVQ_PITCH + k4 * 4];
            if (c0 + cr < size) rvq_cp_async16(rvq_smem_u32(dst), cb + (size_t)(c0 + cr) * RVQ_DIM + k4 * 4);
            else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    prefetch(0);
    float4 q[FPW];
#pragma unroll
    for (int f = 0; f < FPW; ++f) {                           // residual rows of this warp's 4 frames (lane owns 4 dims)
        const long long fr = f0 + warp * FPW + f;
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
        if (fr < frames) r = *reinterpret_cast<const float4*>(z + fr * RVQ_DIM + lane * 4);
        *reinterpret_cast<float4*>(&R[(warp * FPW + f) * RVQ_PITCH + lane * 4]) = r;
        q[f] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    rvq_cluster_sync();                                     // every CTA of the cluster is resident before any DSMEM store

    for (int s = 0; s < n; ++s) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                                    // E[s & 1] complete; every warp is done with E[(s + 1) & 1]
        if (s + 1 < n) prefetch(s + 1);
        const float* E = E0 + (size_t)(s & 1) * RVQ_CT * RVQ_PITCH;
        const float* ees = ee + (size_t)s * size;

        float xx[FPW];
#pragma unroll
        for (int f = 0; f < FPW; ++f) {
            const float4 v = *reinterpret_cast<const float4*>(&R[(warp * FPW + f) * RVQ_PITCH + lane * 4]);
            float p = __fmul_rn(v.x, v.x);
            p = fmaf(v.y, v.y, p); p = fmaf(v.z, v.z, p); p = fmaf(v.w, v.w, p);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
            xx[f] = drop_xx ? 0.f : p;
        }
        float dot[FPW][4];
#pragma unroll
        for (int f = 0; f < FPW; ++f)
#pragma unroll
            for (int j = 0; j < 4; ++j) dot[f][j] = 0.f;
#pragma unroll 4
        for (int k4 = 0; k4 < RVQ_DIM / 4; ++k4) {
            float4 r4[FPW], e4[4];
#pragma unroll
            for (int f = 0; f < FPW; ++f) r4[f] = *reinterpret_cast<const float4*>(&R[(warp * FPW + f) * RVQ_PITCH + k4 * 4]);
#pragma unroll
            for (int j = 0; j < 4; ++j) e4[j] = *reinterpret_cast<const float4*>(&E[(lane + 32 * j) * RVQ_PITCH + k4 * 4]);
#pragma unroll
            for (int f = 0; f < FPW; ++f)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float d = dot[f][j];
                    d = fmaf(r4[f].x, e4[j].x, d);
                    d = fmaf(r4[f].y, e4[j].y, d);
                    d = fmaf(r4[f].z, e4[j].z, d);
                    d = fmaf(r4[f].w, e4[j].w, d);
                    dot[f][j] = d;
                }
        }
        float best[FPW];
        int besti[FPW];
#pragma unroll
        for (int f = 0; f < FPW; ++f) { best[f] = -INFINITY; besti[f] = 0x7fffffff; }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int code = c0 + lane + 32 * j;
            if (code < size) {
                const float e2 = ees[code];
#pragma unroll
                for (int f = 0; f < FPW; ++f) {
                    const float d = -__fadd_rn(__fsub_rn(xx[f], __fmul_rn(2.f, dot[f][j])), e2);
                    if (d > best[f]) { best[f] = d; besti[f] = code; }
                }
            }
        }
        RvqCand* cs = cand + (size_t)(s & 1) * tiles * FT;   // this stage's mailbox: [tiles][FT]
#pragma unroll
        for (int f = 0; f < FPW; ++f) {
            float bd = best[f];
            int bi = besti[f];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float od = __shfl_xor_sync(0xffffffffu, bd, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (od > bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
            }
            // lane t < tiles delivers this CTA's candidate for the frame into CTA t's mailbox
            if (lane < tiles) {
                const uint32_t local = rvq_smem_u32(&cs[tile * FT + warp * FPW + f]);
                uint32_t remote;
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(lane));
                asm volatile("st.shared::cluster.v2.b32 [%0], {%1, %2};" ::"r"(remote), "r"(__float_as_uint(bd)), "r"(bi) : "memory");
            }
        }
        rvq_cluster_sync();                                 // all candidates of stage s have landed everywhere
#pragma unroll
        for (int f = 0; f < FPW; ++f) {
            const long long fr = f0 + warp * FPW + f;
            float bd = -INFINITY;
            int bi = 0x7fffffff;
            for (int t = 0; t < tiles; ++t) {               // ascending code ranges: strict > keeps the first maximum
                const RvqCand c = cs[t * FT + warp * FPW + f];
                if (c.d > bd || (c.d == bd && c.i < bi)) { bd = c.d; bi = c.i; }
            }
            if (bi < 0 || bi >= size) bi = 0;               // all-NaN row, as in the one-kernel search
            if (fr < frames) {
                if (tile == 0 && lane == 0) idx[(size_t)s * frames + fr] = bi;
                const float4 e = *reinterpret_cast<const float4*>(codebooks + ((size_t)s * size + bi) * RVQ_DIM + lane * 4);
                float4 r = *reinterpret_cast<const float4*>(&R[(warp * FPW + f) * RVQ_PITCH + lane * 4]);
                r.x = __fsub_rn(r.x, e.x); r.y = __fsub_rn(r.y, e.y); r.z = __fsub_rn(r.z, e.z); r.w = __fsub_rn(r.w, e.w);
                q[f].x = __fadd_rn(q[f].x, e.x); q[f].y = __fadd_rn(q[f].y, e.y);
                q[f].z = __fadd_rn(q[f].z, e.z); q[f].w = __fadd_rn(q[f].w, e.w);
                *reinterpret_cast<float4*>(&R[(warp * FPW + f) * RVQ_PITCH + lane * 4]) = r;
            }
        }
        __syncwarp();                                       // this warp's R rows are rewritten before it reads them again
    }
    if (tile == 0 && qsum) {
#pragma unroll
        for (int f = 0; f < FPW; ++f) {
            const long long fr = f0 + warp * FPW + f;
            if (fr < frames) *reinterpret_cast<float4*>(qsum + fr * RVQ_DIM + lane * 4) = q[f];
        }
    }
}

static size_t rvq_cluster_smem(int tiles, int ft) {
    return (size_t)(ft + 2 * RVQ_CT) * RVQ_PITCH * sizeof(float) + (size_t)2 * tiles * ft * sizeof(RvqCand);
}

// HILCODEC_RVQ_CLUSTER=0 keeps the per-stage launches (A/B knob)
bool rvq_cluster_usable(int size, int dim, long long frames) {
    static const bool on = [] { const char* e = std::getenv("HILCODEC_RVQ_CLUSTER"); return !(e && e[0] == '0'); }();
    const int tiles = (size + RVQ_CT - 1) / RVQ_CT;
    return on && dim == RVQ_DIM && frames > 0 && frames <= RVQ_SPLIT_MAX_FRAMES && tiles >= 1 && tiles <= 8;
}

template <int FPW>
static cudaError_t launch_cluster_fpw(const float* z, const float* codebooks, const float* ee, int size, int tiles,
                                      long long frames, int n, int64_t* idx, float* qsum, bool drop_xx, cudaStream_t st) {
    constexpr int FT = 8 * FPW;
    const size_t smem = rvq_cluster_smem(tiles, FT);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(rvq_cluster_kernel<FPW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)rvq_cluster_smem(8, FT));
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute attr{};
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = (unsigned)tiles; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.gridDim = dim3((unsigned)tiles, (unsigned)((frames + FT - 1) / FT));
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, rvq_cluster_kernel<FPW>, z, codebooks, ee, size, tiles, frames, n, idx, qsum,
                              drop_xx ? 1 : 0);
}

cudaError_t launch_rvq_encode_cluster(const float* z, const float* codebooks, const float* ee, int size, int dim,
                                      long long frames, int n, int64_t* idx, float* qsum, bool drop_xx, cudaStream_t st) {
    if (dim != RVQ_DIM) return cudaErrorInvalidValue;
    if (frames == 0 || n == 0) return cudaSuccess;
    const int tiles = (size + RVQ_CT - 1) / RVQ_CT;
    // frames per warp: as few as keeps the launch at <= 16 clusters (128 CTAs of 148 SMs); HILCODEC_RVQ_FPW=<1|2|4> forces
    static const int forced = []() { const char* e = std::getenv("HILCODEC_RVQ_FPW"); return e ? std::atoi(e) : 0; }();
    const int fpw = forced ? forced : (frames <= 128 ? 1 : frames <= 256 ? 2 : 4);
    if (fpw == 1) return launch_cluster_fpw<1>(z, codebooks, ee, size, tiles, frames, n, idx, qsum, drop_xx, st);
    if (fpw == 2) return launch_cluster_fpw<2>(z, codebooks, ee, size, tiles, frames, n, idx, qsum, drop_xx, st);
    return launch_cluster_fpw<4>(z, codebooks, ee, size, tiles, frames, n, idx, qsum, drop_xx, st);
}

bool rvq_split_usable(int size, int dim, long long frames) {
    static const bool on = [] { const char* e = std::getenv("HILCODEC_RVQ_SPLIT"); return !(e && e[0] == '0'); }();
    return on && dim == RVQ_DIM && frames > 0 && frames <= RVQ_SPLIT_MAX_FRAMES && (size + RVQ_CT - 1) / RVQ_CT <= RVQ_SPLIT_MAX_TILES;
}

// scratch (rvq_split_scratch_bytes()): 2 * RVQ_SPLIT_MAX_FRAMES * RVQ_SPLIT_MAX_TILES candidates, then the residual / sum rows
static size_t rvq_split_cand_bytes() { return (size_t)2 * RVQ_SPLIT_MAX_FRAMES * RVQ_SPLIT_MAX_TILES * sizeof(RvqCand); }
size_t rvq_split_scratch_bytes() { return rvq_split_cand_bytes() + (size_t)4 * RVQ_SPLIT_MAX_FRAMES * RVQ_DIM * sizeof(float); }

cudaError_t launch_rvq_encode_split(const float* z, const float* codebooks, const float* ee, int size, int dim,
                                    long long frames, int n, int64_t* idx, float* qsum, bool drop_xx, void* scratch,
                                    cudaStream_t st) {
    if (dim != RVQ_DIM || !scratch) return cudaErrorInvalidValue;
    if (frames == 0 || n == 0) return cudaSuccess;
    static bool attr_set = false;
    const size_t smem = (size_t)(RVQ_FT + RVQ_CT) * RVQ_PITCH * sizeof(float);
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(rvq_stage_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const int tiles = (size + RVQ_CT - 1) / RVQ_CT;
    RvqCand* part[2] = {reinterpret_cast<RvqCand*>(scratch),
                        reinterpret_cast<RvqCand*>(scratch) + (size_t)RVQ_SPLIT_MAX_FRAMES * RVQ_SPLIT_MAX_TILES};
    const unsigned fblocks = (unsigned)((frames + RVQ_FT - 1) / RVQ_FT);
    for (int s = 0; s <= n; ++s) {
        const dim3 grid(s == n ? 1 : tiles, fblocks);
        rvq_stage_kernel<<<grid, 256, smem, st>>>(z, codebooks, ee, size, tiles, frames, s, n, idx, qsum, part[(s + 1) & 1],
                                                  part[s & 1],
                                                  reinterpret_cast<float*>(reinterpret_cast<char*>(scratch) + rvq_split_cand_bytes()),
                                                  drop_xx ? 1 : 0);
        const cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// Dequantizer: q[f][:] = ((0 + E_0[i_0]) + E_1[i_1]) + ...   one thread per 4 dims.
__global__ void rvq_decode_kernel(const int64_t* __restrict__ idx, const float* __restrict__ codebooks, int size, int dim,
                                  long long frames, int n, float* __restrict__ q) {
    const int d4 = dim / 4;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= frames * d4) return;
    const long long fr = i / d4;
    const int k4 = (int)(i - fr * d4);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < n; ++s) {
        long long id = idx[(size_t)s * frames + fr];
        if (id < 0 || id >= size) id = 0;  // F.embedding would raise; the host wrapper validates when asked to
        const float4 e = *reinterpret_cast<const float4*>(codebooks + ((size_t)s * size + id) * dim + k4 * 4);
        a.x = __fadd_rn(a.x, e.x); a.y = __fadd_rn(a.y, e.y); a.z = __fadd_rn(a.z, e.z); a.w = __fadd_rn(a.w, e.w);
    }
    *reinterpret_cast<float4*>(q + fr * dim + k4 * 4) = a;
}

cudaError_t launch_rvq_decode(const int64_t* idx, const float* codebooks, int size, int dim, long long frames, int n,
                              float* q, cudaStream_t st) {
    if (dim & 3) return cudaErrorInvalidValue;
    const long long total = frames * (dim / 4);
    if (total == 0) return cudaSuccess;
    rvq_decode_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(idx, codebooks, size, dim, frames, n, q);
    return cudaGetLastError();
}

}  // namespace hil
