// Runs gemm.cu's FP32 FFMA kernel (source text extracted into gemm_extracted.inc) on the CPU emulation layer.
// argv: LD EPI TM Mp M K B T pre pre_scale has_bias has_res hop x_bs x_rs y_bs y_rs M_out in.bin out.bin
//   LD 0 plain, 1 channel-last input, 2 im2col (STFT); EPI 0 linear, 1 log-magnitude; TM 4 / 6 / 8 (BM = 16 * TM)
// in.bin = A[Kp16*Mp] X[nx] bias[M]? R[ny]? (float32); out.bin = Y[ny].
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "cuda_emu.h"
namespace hil {
#include "gemm_extracted.inc"
}
using namespace hil;

template <int TM, int LD, int EPI> void run(const GemmParams& p) {
    emu_launch((p.N + BN - 1) / BN, p.Mp / (16 * TM), 256, [&] { gemm_kernel<TM, LD, EPI>(p); });
}

int main(int argc, char** argv) {
    if (argc != 21) { std::fprintf(stderr, "bad args %d\n", argc); return 2; }
    int a = 1;
    const int LD = atoi(argv[a++]), EPI = atoi(argv[a++]), TM = atoi(argv[a++]);
    GemmParams p{};
    p.Mp = atoi(argv[a++]); p.M = atoi(argv[a++]); p.K = atoi(argv[a++]);
    const int B = atoi(argv[a++]); p.T = atoi(argv[a++]);
    p.pre = atoi(argv[a++]); p.pre_scale = (float)atof(argv[a++]);
    const int has_bias = atoi(argv[a++]), has_res = atoi(argv[a++]);
    p.hop = atoi(argv[a++]); p.x_bs = atoll(argv[a++]); p.x_rs = atoi(argv[a++]);
    p.y_bs = atoll(argv[a++]); p.y_rs = atoi(argv[a++]); p.M_out = atoi(argv[a++]);
    const char* fin = argv[a++]; const char* fout = argv[a++];
    p.N = (unsigned)B * p.T;
    const int Kp = (p.K + 15) / 16 * 16;
    FILE* f = std::fopen(fin, "rb");
    std::fseek(f, 0, SEEK_END); const long bytes = std::ftell(f); std::fseek(f, 0, SEEK_SET);
    std::vector<float> in(bytes / 4);
    if (std::fread(in.data(), 4, in.size(), f) != in.size()) return 3;
    std::fclose(f);
    const size_t nA = (size_t)Kp * p.Mp, ny = (size_t)B * p.y_bs;
    const size_t nx = in.size() - nA - (has_bias ? p.M : 0) - (has_res ? ny : 0);
    p.A = in.data(); p.X = in.data() + nA;
    p.bias = has_bias ? in.data() + nA + nx : nullptr;
    p.R = has_res ? in.data() + nA + nx + (has_bias ? p.M : 0) : nullptr;
    std::vector<float> y(ny, -12345.f);
    p.Y = y.data();
    bool ok = true;
#define CASE(TT, L, E) if (TM == TT && LD == L && EPI == E) run<TT, L, E>(p); else
    CASE(8, 0, 0) CASE(6, 0, 0) CASE(4, 0, 0) CASE(8, 1, 0) CASE(6, 1, 0) CASE(4, 1, 0) CASE(6, 2, 1) ok = false;
#undef CASE
    if (!ok) return 4;
    f = std::fopen(fout, "wb");
    std::fwrite(y.data(), 4, y.size(), f);
    std::fclose(f);
    return 0;
}
