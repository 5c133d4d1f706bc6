// Runs hilcodec_b200/csrc/gemm_skinny.cu's kernel (source text extracted into skinny_extracted.inc) on the CPU
// emulation layer.  argv: LD EPI Mp M K B T pre pre_scale has_bias has_res hop x_bs x_ks y_bs y_rs M_out in.bin out.bin
// in.bin = A[Kp*Mp] X[nx] bias[M]? R[ny]? (float32); out.bin = Y[ny].
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "cuda_emu.h"
namespace hil {
#include "skinny_extracted.inc"
}
using namespace hil;

template <int LD, int EPI> void run(const SkinnyParams& p) {
    const unsigned gy = p.Mp / SK_ROWS;
    auto go = [&](auto nt, auto kc) {
        constexpr int NT = decltype(nt)::value, KC = decltype(kc)::value;
        SkinnyParams q = p;  // launch_one() in gemm_skinny.cu
        q.cols_per_tile = EPI == SK_DW5 ? NT / p.T * p.T : NT;
        emu_launch((q.N + q.cols_per_tile - 1) / q.cols_per_tile, gy, SK_WARPS * 32, [&] { skinny_kernel<NT, KC, LD, EPI>(q); });
    };
    // same choice as dispatch() in gemm_skinny.cu
    if (p.N <= 8) go(std::integral_constant<int, 8>{}, std::integral_constant<int, 32>{});
    else if (p.N <= 16) go(std::integral_constant<int, 16>{}, std::integral_constant<int, 32>{});
    else if (p.N <= 32) go(std::integral_constant<int, 32>{}, std::integral_constant<int, 16>{});
    else go(std::integral_constant<int, 64>{}, std::integral_constant<int, 16>{});
}

int main(int argc, char** argv) {
    if (argc != 20 && argc != 24) { std::fprintf(stderr, "bad args %d\n", argc); return 2; }
    int a = 1;
    const int LD = atoi(argv[a++]), EPI = atoi(argv[a++]);
    SkinnyParams p{};
    p.Mp = atoi(argv[a++]); p.M = atoi(argv[a++]); p.K = atoi(argv[a++]);
    const int B = atoi(argv[a++]); p.T = atoi(argv[a++]);
    p.pre = atoi(argv[a++]); p.pre_scale = (float)atof(argv[a++]);
    const int has_bias = atoi(argv[a++]), has_res = atoi(argv[a++]);
    p.hop = atoi(argv[a++]); p.x_bs = atoll(argv[a++]); p.x_ks = atoll(argv[a++]);
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
    if (EPI == 2) {
        const int has_dwb = atoi(argv[a++]), has_skip = atoi(argv[a++]);
        p.post = atoi(argv[a++]); p.post_scale = (float)atof(argv[a++]);
        const size_t nc = (size_t)B * p.M * 4;
        const size_t nx = in.size() - nA - (size_t)p.M * 5 - (has_dwb ? p.M : 0) - nc - (has_skip ? ny : 0);
        const float* q = in.data();
        p.A = q; q += nA; p.X = q; q += nx; p.dw_w = q; q += (size_t)p.M * 5;
        p.dw_b = has_dwb ? q : nullptr; q += has_dwb ? p.M : 0;
        p.cache_in = q; q += nc; p.skip = has_skip ? q : nullptr;
        std::vector<float> y(ny, -12345.f), co(nc, -12345.f);
        p.Y = y.data(); p.cache_out = co.data();
        if (LD != 0) return 4;
        run<SK_PLAIN, SK_DW5>(p);
        f = std::fopen(fout, "wb");
        std::fwrite(y.data(), 4, y.size(), f);
        std::fwrite(co.data(), 4, co.size(), f);
        std::fclose(f);
        return 0;
    }
    const size_t nx = in.size() - nA - (has_bias ? p.M : 0) - (has_res ? ny : 0);
    p.A = in.data(); p.X = in.data() + nA;
    p.bias = has_bias ? in.data() + nA + nx : nullptr;
    p.R = has_res ? in.data() + nA + nx + (has_bias ? p.M : 0) : nullptr;
    std::vector<float> y(ny, -12345.f);
    p.Y = y.data();
    if (LD == 0 && EPI == 0) run<SK_PLAIN, SK_LINEAR>(p);
    else if (LD == 1 && EPI == 0) run<SK_CHLAST, SK_LINEAR>(p);
    else if (LD == 2 && EPI == 1) run<SK_IM2COL, SK_LOGMAG>(p);
    else return 4;
    f = std::fopen(fout, "wb");
    std::fwrite(y.data(), 4, y.size(), f);
    std::fclose(f);
    return 0;
}
