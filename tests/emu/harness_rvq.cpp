// Runs rvq.cu's one-kernel search (rvq_encode_kernel<8>: 8 frames per warp, launcher-chosen warps per CTA; `slots`
// stands for 2 x SM count) and its per-stage variant (rvq_stage_kernel) -- source text extracted into
// rvq_extracted.inc -- on the CPU emulation layer.
// The batch variant's decision kernel (rvq_tc_select_kernel) runs on dot products computed here in fp64 and then
// perturbed by +-2e-6 |e||r| (twice the tensor-core GEMM's error bound), so its near-tie re-scoring is what keeps the
// indices equal.
// argv: size frames n drop_xx in.bin out.bin slots;  in.bin = z[frames*128] codebooks[n*size*128] (float32);
// out.bin = idx_mono idx_split idx_tc [n*frames each] (int64) qsum_mono qsum_split qsum_tc [frames*128 each] (float32)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "cuda_emu.h"
namespace hil {
#include "rvq_extracted.inc"
}
using namespace hil;

int main(int argc, char** argv) {
    if (argc != 8) return 2;
    const int slots = atoi(argv[7]);
    const int size = atoi(argv[1]);
    const long long frames = atoll(argv[2]);
    const int n = atoi(argv[3]), drop_xx = atoi(argv[4]);
    std::vector<float> z((size_t)frames * RVQ_DIM), cb((size_t)n * size * RVQ_DIM), ee((size_t)n * size);
    FILE* f = std::fopen(argv[5], "rb");
    if (std::fread(z.data(), 4, z.size(), f) != z.size() || std::fread(cb.data(), 4, cb.size(), f) != cb.size()) return 3;
    std::fclose(f);
    for (size_t r = 0; r < ee.size(); ++r) {  // codebook_norm_kernel
        float s = 0.f;
        for (int k = 0; k < RVQ_DIM; ++k) s = __fadd_rn(s, __fmul_rn(cb[r * RVQ_DIM + k], cb[r * RVQ_DIM + k]));
        ee[r] = s;
    }
    std::vector<int64_t> idx_a((size_t)n * frames, -1), idx_b((size_t)n * frames, -1);
    std::vector<float> q_a(z.size(), -1.f), q_b(z.size(), -1.f);
    const unsigned fblocks = (unsigned)((frames + RVQ_FT - 1) / RVQ_FT);
    {   // launch_rvq_encode_big
        const int warps = rvq_v2_warps(frames, 8, slots), ft = warps * 8;
        emu_launch((unsigned)((frames + ft - 1) / ft), 1, warps * 32, [&] {
            rvq_encode_kernel<8>(z.data(), cb.data(), ee.data(), size, frames, n, idx_a.data(), q_a.data(), drop_xx);
        });
        std::fprintf(stderr, "one-kernel search: %d warps per CTA\n", warps);
    }
    const int tiles = (size + RVQ_CT - 1) / RVQ_CT;
    std::vector<float> rq((size_t)4 * RVQ_SPLIT_MAX_FRAMES * RVQ_DIM, -7.f);   // residual / sum hand-over rows
    std::vector<RvqCand> part[2];
    part[0].resize((size_t)frames * tiles); part[1].resize((size_t)frames * tiles);
    for (int s = 0; s <= n; ++s)   // launch_rvq_encode_split
        emu_launch(s == n ? 1 : tiles, fblocks, 256, [&] {
            rvq_stage_kernel(z.data(), cb.data(), ee.data(), size, tiles, frames, s, n, idx_b.data(), q_b.data(),
                             part[(s + 1) & 1].data(), part[s & 1].data(), rq.data(), drop_xx);
        });
    // run_rvq_tc: k-major residuals / sums, one "GEMM" + one decision launch per stage
    std::vector<int64_t> idx_c((size_t)n * frames, -1);
    std::vector<float> q_c(z.size(), -1.f);
    {
        const long long pitch = (frames + 127) / 128 * 128;
        std::vector<float> Rk((size_t)RVQ_DIM * pitch, 0.f), Qk((size_t)RVQ_DIM * pitch, -3.f), Y((size_t)size * pitch, 0.f);
        for (long long fr = 0; fr < frames; ++fr)
            for (int k = 0; k < RVQ_DIM; ++k) Rk[(size_t)k * pitch + fr] = z[(size_t)fr * RVQ_DIM + k];
        unsigned rescored = 0;
        uint32_t lcg = 12345u;
        for (int s = 0; s < n; ++s) {
            const float* cbs = cb.data() + (size_t)s * size * RVQ_DIM;
            float ee_max = 0.f;
            for (int c = 0; c < size; ++c) ee_max = std::fmax(ee_max, ee[(size_t)s * size + c]);
            for (long long fr = 0; fr < frames; ++fr) {
                double rr = 0.0;
                for (int k = 0; k < RVQ_DIM; ++k) rr += (double)Rk[(size_t)k * pitch + fr] * Rk[(size_t)k * pitch + fr];
                for (int c = 0; c < size; ++c) {
                    double dot = 0.0;
                    for (int k = 0; k < RVQ_DIM; ++k) dot += (double)cbs[(size_t)c * RVQ_DIM + k] * Rk[(size_t)k * pitch + fr];
                    lcg = lcg * 1664525u + 1013904223u;
                    const double noise = ((double)(lcg >> 8) / 8388608.0 - 1.0) * 2e-6 *
                                         std::sqrt(rr * (double)ee[(size_t)s * size + c]);
                    Y[((size_t)(fr / RVQ_TC_TILE) * size + c) * RVQ_TC_TILE + fr % RVQ_TC_TILE] = (float)(dot + noise);
                }
            }
            emu_launch((unsigned)((frames + 31) / 32), 1, 256, [&] {
                rvq_tc_select_kernel(Y.data(), Rk.data(), Qk.data(), cbs, ee.data() + (size_t)s * size, size, pitch, frames,
                                     s == 0, idx_c.data() + (size_t)s * frames, ee_max, drop_xx, &rescored);
            });
        }
        for (long long fr = 0; fr < frames; ++fr)   // kmajor_to_rows_kernel
            for (int k = 0; k < RVQ_DIM; ++k) q_c[(size_t)fr * RVQ_DIM + k] = Qk[(size_t)k * pitch + fr];
        std::fprintf(stderr, "tensor-core variant: %u of %lld decisions re-scored\n", rescored, (long long)n * frames);
    }
    f = std::fopen(argv[6], "wb");
    std::fwrite(idx_a.data(), 8, idx_a.size(), f);
    std::fwrite(idx_b.data(), 8, idx_b.size(), f);
    std::fwrite(idx_c.data(), 8, idx_c.size(), f);
    std::fwrite(q_a.data(), 4, q_a.size(), f);
    std::fwrite(q_b.data(), 4, q_b.size(), f);
    std::fwrite(q_c.data(), 4, q_c.size(), f);
    std::fclose(f);
    return 0;
}
