// Runs the small glue kernels of conv.cu / bitpack.cu (source text extracted into misc_extracted.inc) on the CPU
// emulation layer with the launch geometry of their launch_* functions.
//   pack     frames n bits in.bin out.bin            in = idx[n*frames] int64; out = packed bytes, then idx unpacked again (int64)
//   convpre  B C T in.bin out.bin                    in = win[B*(T+4)] w[C*5] bias[C]; out = y[B*C*Tp]
//   convpost B C T pre pre_scale in.bin out.bin      in = x[B*C*Tp] cache[B*C*4] w[C*5] bias[1]; out = y[B*T] cache_out[B*C*4]
//   l2norm   B C F scale in.bin out.bin              in = x[B*C*Fp]; out = z[B*F*C]
//   wavcat   B T P in.bin out.bin                    in = x[B*T] cache[B*P]; out = wav_ext[B*Wp] cache_out[B*P]
//   transp   B C F in.bin out.bin                    in = q[B*F*C]; out = y[B*C*Fp] (chlast_to_ncw), then rows[F*C] of batch 0 back (kmajor_to_rows)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "cuda_emu.h"
namespace hil {
#include "misc_extracted.inc"
}
using namespace hil;

static std::vector<char> slurp(const char* path) {
    FILE* f = std::fopen(path, "rb");
    std::fseek(f, 0, SEEK_END); const long n = std::ftell(f); std::fseek(f, 0, SEEK_SET);
    std::vector<char> b(n + 16);
    if (std::fread(b.data(), 1, n, f) != (size_t)n) std::exit(3);
    std::fclose(f);
    return b;
}
static void dump(const char* path, std::initializer_list<std::pair<const void*, size_t>> parts) {
    FILE* f = std::fopen(path, "wb");
    for (auto& p : parts) std::fwrite(p.first, 1, p.second, f);
    std::fclose(f);
}

int main(int argc, char** argv) {
    const std::string mode = argv[1];
    if (mode == "pack") {
        const long long frames = atoll(argv[2]); const int n = atoi(argv[3]), bits = atoi(argv[4]);
        auto in = slurp(argv[5]);
        const int64_t* idx = reinterpret_cast<const int64_t*>(in.data());
        const int bpf = (n * bits + 7) / 8;
        std::vector<uint8_t> packed((size_t)frames * bpf, 0xAA);
        std::vector<int64_t> back((size_t)n * frames, -1);
        const unsigned g = (unsigned)((frames + 127) / 128);
        emu_launch(g, 1, 128, [&] { pack_indices_kernel(idx, frames, n, bits, bpf, packed.data()); });
        emu_launch(g, 1, 128, [&] { unpack_indices_kernel(packed.data(), frames, n, bits, bpf, back.data()); });
        dump(argv[6], {{packed.data(), packed.size()}, {back.data(), back.size() * 8}});
    } else if (mode == "convpre") {
        const int B = atoi(argv[2]), C = atoi(argv[3]), T = atoi(argv[4]);
        auto in = slurp(argv[5]);
        const float* win = reinterpret_cast<const float*>(in.data());
        const float* w = win + (size_t)B * (T + 4); const float* bias = w + C * 5;
        const int Tp = (T + 3) & ~3;
        std::vector<float> y((size_t)B * C * Tp, -12345.f);
        emu_launch((T + 511) / 512, B, 128, [&] { conv_pre_kernel<5>(win, T + 4, w, bias, y.data(), (long long)C * Tp, Tp, C, T); });
        dump(argv[6], {{y.data(), y.size() * 4}});
    } else if (mode == "convpost") {
        const int B = atoi(argv[2]), C = atoi(argv[3]), T = atoi(argv[4]), pre = atoi(argv[5]);
        const float pre_scale = (float)atof(argv[6]);
        auto in = slurp(argv[7]);
        const int Tp = (T + 3) & ~3;
        const float* x = reinterpret_cast<const float*>(in.data());
        const float* ci = x + (size_t)B * C * Tp; const float* w = ci + (size_t)B * C * 4; const float* bias = w + C * 5;
        std::vector<float> y((size_t)B * T, -12345.f), co((size_t)B * C * 4, -12345.f);
        const int Tq = (T + 7) / 8, th = Tq >= 128 ? 128 : 32;
        emu_launch(max(1, (Tq + th - 1) / th), B, th, [&] {
            conv_post_tanh_kernel<5>(x, (long long)C * Tp, Tp, ci, co.data(), w, bias, y.data(), C, T, pre, pre_scale, 1, nullptr);
        });
        dump(argv[8], {{y.data(), y.size() * 4}, {co.data(), co.size() * 4}});
    } else if (mode == "l2norm") {
        const int B = atoi(argv[2]), C = atoi(argv[3]), Fr = atoi(argv[4]);
        const float scale = (float)atof(argv[5]);
        auto in = slurp(argv[6]);
        const int Fp = (Fr + 3) & ~3;
        std::vector<float> z((size_t)B * Fr * C, -12345.f);
        const long long total = (long long)B * Fr;
        emu_launch((unsigned)((total * 32 + 255) / 256), 1, 256, [&] {
            l2norm_chlast_kernel(reinterpret_cast<const float*>(in.data()), (long long)C * Fp, Fp, z.data(), C, Fr, total, scale, nullptr);
        });
        dump(argv[7], {{z.data(), z.size() * 4}});
    } else if (mode == "wavcat") {
        const int B = atoi(argv[2]), T = atoi(argv[3]), P = atoi(argv[4]);
        auto in = slurp(argv[5]);
        const float* x = reinterpret_cast<const float*>(in.data()); const float* ci = x + (size_t)B * T;
        const int Wp = (P + T + 3) & ~3;
        std::vector<float> ext((size_t)B * Wp, -12345.f), co((size_t)B * P, -12345.f);
        emu_launch(min((P + T + 255) / 256, 1024), B, 256, [&] { wavcat_kernel(x, ci, co.data(), ext.data(), Wp, T, P); });
        dump(argv[6], {{ext.data(), ext.size() * 4}, {co.data(), co.size() * 4}});
    } else if (mode == "transp") {
        const int B = atoi(argv[2]), C = atoi(argv[3]), F = atoi(argv[4]);
        auto in = slurp(argv[5]);
        const float* q = reinterpret_cast<const float*>(in.data());
        const int Fp = (F + 3) & ~3;
        std::vector<float> y((size_t)B * C * Fp, 0.f), rows((size_t)F * C, -1.f);
        // launch_chlast_to_ncw: grid (F/32, C/32, B), block (32, 8)
        emu_launch((F + 31) / 32, (C + 31) / 32, 32, [&] { chlast_to_ncw_kernel(q, y.data(), C, F, (long long)C * Fp, Fp); }, B, 8);
        // launch_kmajor_to_rows on batch 0's [C][Fp] block: grid (F/32, C/32), block (32, 8)
        emu_launch((F + 31) / 32, (C + 31) / 32, 32, [&] { kmajor_to_rows_kernel(y.data(), Fp, rows.data(), C, F); }, 1, 8);
        dump(argv[6], {{y.data(), y.size() * 4}, {rows.data(), rows.size() * 4}});
    } else {
        return 4;
    }
    return 0;
}
