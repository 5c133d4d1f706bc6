// fp16 hi/lo split arithmetic shared by the kind::f16 tensor-core kernels (gemm_h.cu, gemm_rb.cu):
//   x = hi + lo * 2^-11,  hi = fp16_rn(x),  lo = fp16_rn((x - hi) * 2^11)
// and the tcgen05 kind::f16 instruction wrapper.  See gemm_h.cu for the number format.
#pragma once
#include <cuda_fp16.h>

#include "tc_ptx.cuh"

namespace hil {
namespace th {

using namespace tc;

constexpr float LO_SCALE = 2048.f;                // 2^11

// kind::f16, A = B = fp16, D = f32, A K-major, B MN-major (cute::UMMA::InstrDescriptor bit layout)
constexpr uint32_t make_idesc_f16(int m, int n) {
    return (1u << 4) | (0u << 7) | (0u << 10) | (0u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ uint32_t pack_h2(float lo_elem, float hi_elem) {
    const __half2 h = __floats2half2_rn(lo_elem, hi_elem);   // .x (low 16 bits) = first element in memory
    return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_h2(uint32_t u) { return __half22float2(*reinterpret_cast<const __half2*>(&u)); }

// Transform of 4 consecutive activations: optional ELU prologue, then the fp16 hi/lo split, on the packed
// fp32 pipe (FMUL2 / FADD2 / FFMA2: ~7 issue slots per element instead of ~20 with scalar code and elu_fast).
//   ELU(x) = x > 0 ? x : 2^(x log2 e) - 1 with ex2.approx: absolute error <= 2.4e-7, the bound elu_fast already
//   has below -1/16 (its near-zero polynomial only buys relative accuracy for |x| < 1/16, which the GEMM that
//   consumes the value cannot see: the error enters the dot product in absolute terms).
//   x - fp16(x) is exact in fp32 and the 2^11 scaling is a power of two, so the split is bit-identical to the
//   scalar form  lo = fp16_rn((x - hi) * 2^11).
// ELU of 4 values on the packed pipe: x > 0 ? x : 2^(x log2 e) - 1 (the function split4<PRE_ELU> applies)
__device__ __forceinline__ void elu4(float& r0, float& r1, float& r2, float& r3) {
    const f32x2 l2e = pk2(1.4426950408889634f, 1.4426950408889634f), m1 = pk2(-1.f, -1.f);
    float t0, t1, t2, t3;
    upk2(fmul2(pk2(r0, r1), l2e), t0, t1);
    upk2(fmul2(pk2(r2, r3), l2e), t2, t3);
    float e0, e1, e2, e3;
    upk2(fadd2(pk2(ex2_approx(t0), ex2_approx(t1)), m1), e0, e1);
    upk2(fadd2(pk2(ex2_approx(t2), ex2_approx(t3)), m1), e2, e3);
    r0 = r0 > 0.f ? r0 : e0;
    r1 = r1 > 0.f ? r1 : e1;
    r2 = r2 > 0.f ? r2 : e2;
    r3 = r3 > 0.f ? r3 : e3;
}

template <int kPre, bool kPoly = false>
__device__ __forceinline__ void split4(float4 x, float s, uint32_t& h01, uint32_t& h23, uint32_t& l01, uint32_t& l23) {
    float r0 = x.x, r1 = x.y, r2 = x.z, r3 = x.w;
    if (kPre != PRE_NONE && kPoly) {   // HILCODEC_ELU_POLY=1: elu_fast (polynomial near zero), for A/B accuracy checks
        r0 = elu_fast(r0 * s); r1 = elu_fast(r1 * s); r2 = elu_fast(r2 * s); r3 = elu_fast(r3 * s);
    } else if (kPre != PRE_NONE) {
        f32x2 a = pk2(r0, r1), b = pk2(r2, r3);
        if (kPre == PRE_SCALE_ELU) {
            const f32x2 s2 = pk2(s, s);
            a = fmul2(a, s2);
            b = fmul2(b, s2);
            upk2(a, r0, r1);
            upk2(b, r2, r3);
        }
        elu4(r0, r1, r2, r3);
    }
    h01 = pack_h2(r0, r1);
    h23 = pack_h2(r2, r3);
    const float2 f01 = unpack_h2(h01), f23 = unpack_h2(h23);
    const f32x2 m1 = pk2(-1.f, -1.f), sc = pk2(LO_SCALE, LO_SCALE);
    float d0, d1, d2, d3;
    upk2(fmul2(ffma2(pk2(f01.x, f01.y), m1, pk2(r0, r1)), sc), d0, d1);
    upk2(fmul2(ffma2(pk2(f23.x, f23.y), m1, pk2(r2, r3)), sc), d2, d3);
    l01 = pack_h2(d0, d1);
    l23 = pack_h2(d2, d3);
}

}  // namespace th
}  // namespace hil
