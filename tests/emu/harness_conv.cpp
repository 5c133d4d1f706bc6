// Runs conv.cu's causal depthwise kernels (source text extracted into conv_extracted.inc) on the CPU emulation layer
// with the launch geometry of launch_dwconv / launch_dwconv_transpose (dw_block / dw_grid, extracted too: short rows run
// many per CTA with a two-dimensional block).
// argv: mode K S B C T pre pre_scale has_bias has_skip post post_scale in.bin out.bin
//   mode 0 dwconv_kernel<K,S>   1 dwconv5_kernel   2 dwconv_strided4_kernel<K,S>   3 dwconvT_kernel<S> (K = 2S)
// rows are pitched to a multiple of 4 floats.  in.bin = x[B*C*Tp] cache[B*C*P] w[C*K] bias[C]? skip[B*C*Top]?
// out.bin = y[B*C*Top] cache_out[B*C*P]        (mode 3: P = 1, no bias / skip / post, T_out = S*T)
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "cuda_emu.h"
namespace hil {
#include "conv_extracted.inc"
}
using namespace hil;

int main(int argc, char** argv) {
    if (argc != 15) return 2;
    int a = 1;
    const int mode = atoi(argv[a++]), K = atoi(argv[a++]), S = atoi(argv[a++]), B = atoi(argv[a++]), C = atoi(argv[a++]);
    const int T = atoi(argv[a++]), pre = atoi(argv[a++]);
    const float pre_scale = (float)atof(argv[a++]);
    const int has_bias = atoi(argv[a++]), has_skip = atoi(argv[a++]), post = atoi(argv[a++]);
    const float post_scale = (float)atof(argv[a++]);
    const char* fin = argv[a++]; const char* fout = argv[a++];
    const int P = mode == 3 ? 1 : K - S;
    const int T_out = mode == 3 ? S * T : (P + T - K) / S + 1;
    const int Tp = (T + 3) & ~3, Top = (T_out + 3) & ~3;
    const size_t nx = (size_t)B * C * Tp, nc = (size_t)B * C * P, nw = (size_t)C * K, ny = (size_t)B * C * Top;
    std::vector<float> in(nx + nc + nw + (has_bias ? C : 0) + (has_skip ? ny : 0));
    FILE* f = std::fopen(fin, "rb");
    if (std::fread(in.data(), 4, in.size(), f) != in.size()) return 3;
    std::fclose(f);
    const float* x = in.data(); const float* ci = x + nx; const float* w = ci + nc;
    const float* bias = has_bias ? w + nw : nullptr;
    const float* skip = has_skip ? w + nw + (has_bias ? C : 0) : nullptr;
    std::vector<float> y(ny, -12345.f), co(nc, -12345.f);
    const long long x_bs = (long long)C * Tp, y_bs = (long long)C * Top;
    // geometry of the launchers: block = dw_block(time threads, min threads along time), grid = dw_grid(...)
    auto launch = [&](int time_threads, int min_tx, auto kernel) {
        const dim3 block = dw_block(time_threads, min_tx);
        dim3 grid = dw_grid(time_threads, block, C, B);
        if (grid.x < 1) grid.x = 1;
        std::fprintf(stderr, "block (%u, %u) grid (%u, %u, %u)\n", block.x, block.y, grid.x, grid.y, grid.z);
        emu_launch(grid.x, grid.y, block.x, kernel, grid.z, block.y);
    };
#define ARGS x, x_bs, Tp, ci, co.data(), w, bias, skip, y.data(), y_bs, Top, C, T
    if (mode == 1) {
        launch((T + 3) / 4, 4, [&] { dwconv5_kernel(ARGS, pre, pre_scale, post, post_scale); });
    } else if (mode == 0 || mode == 2) {
        const int n = mode == 2 ? (T_out + 3) / 4 : T_out;
#define RUN(KK, SS)                                                                                                      \
    if (K == KK && S == SS) {                                                                                            \
        if (mode == 2) launch(n, P, [&] { dwconv_strided4_kernel<KK, SS>(ARGS, T_out, pre, pre_scale, post, post_scale); }); \
        else launch(n, P, [&] { dwconv_kernel<KK, SS>(ARGS, T_out, pre, pre_scale, post, post_scale); });               \
    }
        RUN(4, 2) RUN(8, 4) RUN(10, 5) RUN(16, 8)
        if (K == 5 && S == 1 && mode == 0)
            launch(n, P, [&] { dwconv_kernel<5, 1>(ARGS, T_out, pre, pre_scale, post, post_scale); });
#undef RUN
    } else if (mode == 3) {
#define RUNT(SS) \
    if (S == SS) launch((T + 3) / 4, 1, [&] { dwconvT_kernel<SS>(x, x_bs, Tp, ci, co.data(), w, y.data(), y_bs, Top, C, T, pre, pre_scale, 1); });
        RUNT(2) RUNT(4) RUNT(5) RUNT(8)
#undef RUNT
    } else {
        return 4;
    }
    f = std::fopen(fout, "wb");
    std::fwrite(y.data(), 4, y.size(), f);
    std::fwrite(co.data(), 4, co.size(), f);
    std::fclose(f);
    return 0;
}
