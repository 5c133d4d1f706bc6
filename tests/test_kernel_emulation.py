"""The SOURCE of the small FP32 CUDA kernels, compiled for the CPU against a thread-per-CUDA-thread emulation
layer (tests/emu/cuda_emu.h) and checked against numpy / the oracle.

These kernels (`gemm_skinny.cu`, the per-stage RVQ kernel in `rvq.cu`) were written when no GPU time was left, so
this is the only execution their logic has had: it verifies indexing, the fixed-order reductions and the arithmetic
of the very text nvcc compiles, not performance and not GPU memory-model behaviour.  The tcgen05 kernels cannot be
emulated this way and are covered by the `-m gpu` tests only."""
import os
import re
import shutil
import subprocess

import numpy as np
import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "hilcodec_b200", "csrc")
EMU = os.path.join(ROOT, "tests", "emu")

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")


def _function(src: str, start: int) -> str:
    """Text from `start` to the closing brace of the first top-level block after it."""
    i = src.index("{", start)
    depth = 0
    for j in range(i, len(src)):
        depth += src[j] == "{"
        depth -= src[j] == "}"
        if depth == 0:
            return src[start:j + 1]
    raise ValueError("unbalanced braces")


def _common_bits() -> str:
    src = open(os.path.join(CSRC, "common.cuh")).read()
    out = [re.search(r"enum Pre \{[^}]*\};", src).group(0)]
    for name in ("elu1", "apply_pre"):
        m = re.search(rf"__device__ __forceinline__ float {name}\(", src)
        out.append(_function(src, m.start()))
    return "\n".join(out)


def _build(tmp, name, inc_name, inc_text, harness):
    with open(os.path.join(tmp, inc_name), "w") as f:
        f.write(inc_text)
    exe = os.path.join(tmp, name)
    cmd = ["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-pthread", "-I", EMU, "-I", tmp,
           os.path.join(EMU, harness), "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return exe


@pytest.fixture(scope="module")
def skinny_exe(tmp_path_factory):
    src = open(os.path.join(CSRC, "gemm_skinny.cu")).read()
    a = src.index("enum { SK_PLAIN")
    k = src.index("template <int NT, int KC, int LD, int EPI>\n__global__")
    text = src[a:k] + _function(src, k)
    text = text.replace("extern __shared__ __align__(16) float sk_smem[];", "float* sk_smem = g_dyn_smem;")
    assert "g_dyn_smem" in text
    # apply_act_fast uses ex2.approx through inline PTX: stand-in with the same meaning (differs by <= 2.4e-7)
    act = ("inline float apply_act_fast(float x, int mode, float s) {\n"
           "    if (mode == PRE_NONE) return x;\n    if (mode == PRE_SCALE_ELU) x = x * s;\n"
           "    return x > 0.f ? x : expm1f(x);\n}\n")
    return _build(str(tmp_path_factory.mktemp("emu_skinny")), "skinny", "skinny_extracted.inc",
                  _common_bits() + "\n" + act + text, "harness_skinny.cpp")


def _pack_kmajor(w, Mp):
    """[M][K] -> k-major [Kp16][Mp], zero padded (the layout hil_model_finalize gives gemm.cu / gemm_skinny.cu)."""
    M, K = w.shape
    Kp = (K + 15) // 16 * 16
    a = np.zeros((Kp, Mp), np.float32)
    a[:K, :M] = w.T
    return a


def _run(exe, tmp, LD, EPI, A, Mp, M, K, B, T, pre, pre_scale, X, bias, R, hop, x_bs, x_ks, y_bs, y_rs, M_out):
    fin, fout = os.path.join(tmp, "in.bin"), os.path.join(tmp, "out.bin")
    parts = [A.ravel(), X.ravel()] + ([bias.ravel()] if bias is not None else []) + ([R.ravel()] if R is not None else [])
    np.concatenate(parts).astype(np.float32).tofile(fin)
    args = [LD, EPI, Mp, M, K, B, T, pre, pre_scale, int(bias is not None), int(R is not None), hop, x_bs, x_ks, y_bs, y_rs,
            M_out, fin, fout]
    r = subprocess.run([exe] + [str(v) for v in args], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.returncode, r.stderr[-2000:])
    return np.fromfile(fout, np.float32)


@pytest.mark.parametrize("B,M,K,T,pre,bias,res", [
    (1, 96, 192, 8, 2, True, True),      # NT = 8, K split over the warps, pre-scale + ELU, bias, residual
    (1, 64, 33, 1, 0, True, True),       # one column, odd K (SpecBlock 1x1), two of the eight warps busy
    (3, 128, 100, 5, 1, False, False),   # N = 15 -> NT = 16, columns of three streams
    (2, 64, 64, 13, 0, True, False),     # N = 26 -> NT = 32, KC = 16
    (1, 96, 40, 40, 1, False, True),     # N = 40 -> NT = 64
    (2, 32, 300, 40, 2, True, True),     # N = 80 -> two column tiles, second one ragged
    (1, 64, 1536, 8, 1, True, False),    # the widest layer: all eight warps, six chunks of 32 k each
    (1, 32, 256, 4, 0, False, False),    # K = 8 warps x one chunk
])
def test_skinny_linear_source_on_cpu(skinny_exe, tmp_path, B, M, K, T, pre, bias, res):
    g = torch.Generator().manual_seed(M + K + T)
    x = torch.randn(B, K, T, generator=g)
    w = torch.randn(M, K, generator=g) / K ** 0.5
    b = torch.randn(M, generator=g) if bias else None
    r = torch.randn(B, M, T, generator=g) if res else None
    xin = x if pre == 0 else F.elu(x * 0.8660254 if pre == 2 else x)
    ref = F.conv1d(xin.double(), w.double()[:, :, None], b.double() if bias else None)
    if res:
        ref = ref + r.double()
    Mp = (M + 31) // 32 * 32
    y = _run(skinny_exe, str(tmp_path), 0, 0, _pack_kmajor(w.numpy(), Mp), Mp, M, K, B, T, pre, 0.8660254, x.numpy(),
             b.numpy() if bias else None, r.numpy() if res else None, 0, K * T, T, M * T, T, M)
    assert np.abs(y.reshape(B, M, T) - ref.numpy()).max() < 2e-5


@pytest.mark.parametrize("B,C,K,T,pre,dwb,skip,post", [
    (1, 96, 96, 8, 1, True, True, 0),     # one stream, one hop at 75*8 Hz: ResBlock tail (skip), NT = 8
    (1, 64, 128, 1, 2, True, False, 1),   # one frame per call: the window is cache + 1 sample; store-side ELU
    (5, 32, 64, 3, 1, False, False, 0),   # five streams of 3 columns -> NT = 16 holds 5 whole streams
    (3, 32, 48, 40, 0, True, True, 2),    # NT = 64 holds one 40-column stream per tile -> 3 column tiles
    (9, 64, 32, 8, 1, True, True, 0),     # NT = 64: 8 streams per tile, second tile has one
])
def test_skinny_fused_dws_source_on_cpu(skinny_exe, tmp_path, B, C, K, T, pre, dwb, skip, post):
    """DWSBlock for a short chunk in one launch: act -> 1x1 -> causal depthwise k5 (+cache) + bias + skip -> act."""
    g = torch.Generator().manual_seed(C + K + T)
    x = torch.randn(B, K, T, generator=g)
    w = torch.randn(C, K, generator=g) / K ** 0.5
    dw_w = torch.randn(C, 1, 5, generator=g) / 5 ** 0.5
    dw_b = torch.randn(C, generator=g) if dwb else None
    cache = torch.randn(B, C, 4, generator=g)
    sk = torch.randn(B, C, T, generator=g) if skip else None
    xin = x if pre == 0 else F.elu(x * 0.8660254 if pre == 2 else x)
    v = F.conv1d(xin.double(), w.double()[:, :, None])
    cat = torch.cat((cache.double(), v), 2)
    ref = F.conv1d(cat, dw_w.double(), dw_b.double() if dwb else None, groups=C)
    if skip:
        ref = ref + sk.double()
    if post:
        ref = F.elu(ref * 0.7 if post == 2 else ref)
    fin, fout = os.path.join(str(tmp_path), "in.bin"), os.path.join(str(tmp_path), "out.bin")
    Mp = (C + 31) // 32 * 32
    parts = [_pack_kmajor(w.numpy(), Mp).ravel(), x.numpy().ravel(), dw_w.numpy().ravel()]
    parts += ([dw_b.numpy()] if dwb else []) + [cache.numpy().ravel()] + ([sk.numpy().ravel()] if skip else [])
    np.concatenate(parts).astype(np.float32).tofile(fin)
    args = [0, 2, Mp, C, K, B, T, pre, 0.8660254, 0, 0, 0, K * T, T, C * T, T, C, fin, fout, int(dwb), int(skip), post, 0.7]
    r = subprocess.run([skinny_exe] + [str(a) for a in args], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.returncode, r.stderr[-2000:])
    out = np.fromfile(fout, np.float32)
    y, co = out[:B * C * T].reshape(B, C, T), out[B * C * T:].reshape(B, C, 4)
    assert np.abs(y - ref.numpy()).max() < 3e-5
    assert np.abs(co - cat[:, :, -4:].numpy()).max() < 3e-5


def test_skinny_channel_last_source_on_cpu(skinny_exe, tmp_path):
    """Decoder input: q [B, F, 128] channel-last -> [B, M, F]."""
    g = torch.Generator().manual_seed(7)
    B, Fr, K, M = 2, 3, 128, 96
    q = torch.randn(B, Fr, K, generator=g)
    w = torch.randn(M, K, generator=g) / K ** 0.5
    ref = F.conv1d(q.transpose(1, 2).double(), w.double()[:, :, None])
    y = _run(skinny_exe, str(tmp_path), 1, 0, _pack_kmajor(w.numpy(), M), M, M, K, B, Fr, 0, 1.0, q.numpy(), None, None,
             0, 0, 1, M * Fr, Fr, M)
    assert np.abs(y.reshape(B, M, Fr) - ref.numpy()).max() < 2e-5


@pytest.mark.parametrize("B,n_fft,hop,T", [(2, 64, 40, 3), (1, 128, 320, 1)])
def test_skinny_stft_source_on_cpu(skinny_exe, tmp_path, B, n_fft, hop, T):
    """DFT-as-conv + magnitude + clamp + log; rows interleaved (cos_f, sin_f) as hil_model_finalize packs them."""
    from hilcodec_b200.weights import dft_basis
    g = torch.Generator().manual_seed(n_fft)
    L = (T - 1) * hop + n_fft
    wav = 0.1 * torch.randn(B, 1, L, generator=g)
    wav[0, 0, : L // 2] = 0.0
    basis = torch.from_numpy(dft_basis(n_fft))  # [2F,1,N] = [cos; sin]
    Fr = n_fft // 2 + 1
    s = F.conv1d(wav.double(), basis.double(), None, stride=hop).view(B, 2, Fr, T)
    ref = s.square().sum(1).sqrt().clamp_min(1e-5).log()
    inter = torch.stack((basis[:Fr, 0], basis[Fr:, 0]), 1).reshape(2 * Fr, n_fft)  # rows (cos_f, sin_f)
    Mp = (2 * Fr + 95) // 96 * 96
    y = _run(skinny_exe, str(tmp_path), 2, 1, _pack_kmajor(inter.numpy(), Mp), Mp, 2 * Fr, n_fft, B, T, 0, 1.0, wav.numpy(),
             None, None, hop, L, 1, Fr * T, T, Fr)
    err = np.abs(y.reshape(B, Fr, T) - ref.numpy())
    assert np.median(err) < 1e-5 and err.max() < 1e-2  # the log amplifies fp32 noise where the magnitude cancels


@pytest.fixture(scope="module")
def rvq_exe(tmp_path_factory):
    src = open(os.path.join(CSRC, "rvq.cu")).read()
    a = src.index("constexpr int RVQ_SPLIT_MAX_FRAMES")
    b = src.index("constexpr int RVQ_PITCH")
    consts = src[a:src.index("\n", b) + 1]
    mono = _function(src, src.index("template <int FPW>\n__global__ void __launch_bounds__(288, 2)\nrvq_encode_kernel("))
    stage = _function(src, src.index("__global__ void __launch_bounds__(256, 2)\nrvq_stage_kernel("))
    warps = _function(src, src.index("int rvq_v2_warps("))
    select = "constexpr int RVQ_TC_TILE = 128;\nconstexpr int RVQ_TC_CPW = 256;\n" + _function(
        src, src.index("__global__ void __launch_bounds__(256, 5)\nrvq_tc_select_kernel("))
    text = consts + mono + "\nstruct RvqCand { float d; int i; };\n" + stage + "\n" + warps + "\n" + select
    assert text.count("extern __shared__ __align__(16) float smem[];") == 2
    text = text.replace("extern __shared__ __align__(16) float smem[];", "float* smem = g_dyn_smem;")
    return _build(str(tmp_path_factory.mktemp("emu_rvq")), "rvq", "rvq_extracted.inc", text, "harness_rvq.cpp")


@pytest.mark.parametrize("size,frames,n,drop_xx,slots", [(1024, 5, 3, 0, 296), (256, 40, 4, 1, 2), (200, 33, 2, 0, 1),
                                                          (128, 100, 2, 0, 1)])
def test_rvq_kernels_source_on_cpu(rvq_exe, tmp_path, size, frames, n, drop_xx, slots):
    """The one-kernel search (1, 3, 5 and 8 warps per CTA in these cases) against the oracle; the per-stage variant and
    the tensor-core variant's decision kernel bit-identical to it."""
    from oracle import hilcodec_oracle as O

    g = torch.Generator().manual_seed(size + frames)
    z = F.normalize(torch.randn(1, frames, 128, generator=g), dim=2) * 128 ** 0.5
    cbs = [torch.randn(size, 128, generator=g) * 0.7 ** i for i in range(n)]
    cbs[0][7] = cbs[0][3]  # an exact tie: the first index must win
    if size > 130:
        cbs[0][130] = cbs[0][3]  # ... also across code tiles (codes 0-127 | 128-255)
    z[0, 0] = cbs[0][3]
    fin, fout = os.path.join(str(tmp_path), "in.bin"), os.path.join(str(tmp_path), "out.bin")
    np.concatenate([z.numpy().ravel()] + [c.numpy().ravel() for c in cbs]).astype(np.float32).tofile(fin)
    r = subprocess.run([rvq_exe, str(size), str(frames), str(n), str(drop_xx), fin, fout, str(slots)], capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    raw = np.fromfile(fout, np.uint8)
    ni = n * frames * 8
    idx_a = raw[:ni].view(np.int64).reshape(n, frames)
    idx_b = raw[ni:2 * ni].view(np.int64).reshape(n, frames)
    idx_c = raw[2 * ni:3 * ni].view(np.int64).reshape(n, frames)
    q = raw[3 * ni:].view(np.float32).reshape(3, frames, 128)
    assert np.array_equal(idx_a, idx_b) and np.array_equal(q[0], q[1])  # split == one-kernel, bit for bit
    # the batch variant's decision kernel on perturbed dot products (stand-in for the tensor-core GEMM): also bit for bit
    assert np.array_equal(idx_a, idx_c) and np.array_equal(q[0], q[2])
    assert "re-scored" in r.stderr and int(r.stderr.split("tensor-core variant: ")[1].split()[0]) >= 1
    assert not (idx_a[0] == 7).any() and not (idx_a[0] == 130).any()
    assert idx_a[0, 0] == 3  # frame 0 IS code 3 of stage 0 (set below): its two exact copies must lose
    p = {f"quantizer.layers.{i}.embed": c for i, c in enumerate(cbs)}
    cfg = O.CodecConfig(num_quantizers=n, codebook_size=size)
    if drop_xx:
        q_ref, _, idx_ref = O.rvq_forward_train(cfg, p, z.transpose(1, 2), n)
        idx_ref, q_ref = idx_ref.permute(1, 0, 2), q_ref.transpose(1, 2)
    else:
        idx_ref = O.rvq_encode(cfg, p, z, n)
        q_ref = O.rvq_decode(cfg, p, idx_ref, n)
    same = idx_a == idx_ref[:, 0].numpy()
    if not same.all():  # near-tie policy of the GPU tests
        _, gaps = O.rvq_margins(cfg, p, z, n)
        first = np.argmax(~same, axis=0)
        bad = np.nonzero((~same).any(axis=0))[0]
        assert all(float(gaps[first[f], 0, f]) < 1e-5 for f in bad)
    else:
        assert np.array_equal(q[0], q_ref[0].numpy())


# ------------------------------------------------------------------------------------------ conv.cu
@pytest.fixture(scope="module")
def conv_exe(tmp_path_factory):
    src = open(os.path.join(CSRC, "conv.cu")).read()
    parts = []
    for sig in ("template <int K, int S>\n__global__ void dwconv_kernel(", "__global__ void dwconv5_kernel(",
                "template <int K, int S>\n__global__ void dwconv_strided4_kernel(",
                "template <int S>\n__global__ void dwconvT_kernel(", "static inline dim3 dw_block(", "static inline dim3 dw_grid("):
        parts.append(_function(src, src.index(sig)))
    act = ("inline float apply_act_fast(float x, int mode, float s) {\n"
           "    if (mode == PRE_NONE) return x;\n    if (mode == PRE_SCALE_ELU) x = x * s;\n"
           "    return x > 0.f ? x : expm1f(x);\n}\n")
    return _build(str(tmp_path_factory.mktemp("emu_conv")), "conv", "conv_extracted.inc",
                  _common_bits() + "\n" + act + "\n".join(parts), "harness_conv.cpp")


def _pitched(t):
    """[B,C,T] -> rows padded to a multiple of 4 floats (the library's activation layout)."""
    B, C, T = t.shape
    Tp = (T + 3) // 4 * 4
    out = np.zeros((B, C, Tp), np.float32)
    out[:, :, :T] = t.numpy()
    return out


@pytest.mark.parametrize("mode,K,S,B,C,T,pre,bias,skip,post", [
    (1, 5, 1, 2, 3, 37, 1, True, True, 0),     # dwconv5: 4 outputs per thread, ragged last group, ELU prologue, skip
    (1, 5, 1, 1, 2, 1, 0, True, False, 1),     # one frame per call: the window is the cache + 1 sample; store-side ELU
    (1, 5, 1, 1, 2, 600, 2, False, False, 2),  # 150 groups -> 128-thread CTAs, two of them
    (0, 5, 1, 2, 2, 3, 0, True, False, 0),     # generic kernel, chunk shorter than the cache
    (0, 16, 8, 1, 3, 8, 0, True, False, 0),    # encoder downsampling, one frame per call (k = 2r, s = r)
    (0, 10, 5, 2, 2, 40, 1, True, True, 0),
    (2, 4, 2, 2, 3, 320, 0, True, False, 0),   # vectorised strided kernel
    (2, 8, 4, 1, 2, 72, 0, True, False, 0),    # 18 outputs: ragged last group of 4
    (2, 10, 5, 1, 2, 600, 0, True, True, 1),   # stride 5: unaligned history
    (2, 16, 8, 1, 2, 600, 0, False, False, 0),
    # short rows (streaming): many rows per CTA, two-dimensional blocks (dw_block)
    (1, 5, 1, 3, 70, 8, 1, True, True, 0),      # k5: 2 time threads -> block (4, 32), 70 channels over 3 CTAs
    (1, 5, 1, 2, 40, 1, 0, True, False, 1),     # one sample per row
    (0, 16, 8, 2, 20, 8, 2, True, False, 0),    # stride 8: one output per row, 8 cache values -> block (8, 16)
    (0, 10, 5, 3, 33, 40, 0, False, False, 0),  # 8 outputs per row
])
def test_depthwise_kernels_source_on_cpu(conv_exe, tmp_path, mode, K, S, B, C, T, pre, bias, skip, post):
    """CausalConv1d.forward causal_layers.py:160-165 (depthwise) through the three kernels of conv.cu."""
    g = torch.Generator().manual_seed(K * 100 + T)
    P = K - S
    x = torch.randn(B, C, T, generator=g)
    cache = torch.randn(B, C, P, generator=g)
    w = torch.randn(C, 1, K, generator=g) / K ** 0.5
    b = torch.randn(C, generator=g) if bias else None
    xin = torch.cat((cache, x if pre == 0 else F.elu(x * 0.77 if pre == 2 else x)), 2)
    ref = F.conv1d(xin.double(), w.double(), b.double() if bias else None, stride=S, groups=C)
    T_out = ref.shape[2]
    sk = torch.randn(B, C, T_out, generator=g) if skip else None
    if skip:
        ref = ref + sk.double()
    if post:
        ref = F.elu(ref * 0.6 if post == 2 else ref)
    fin, fout = os.path.join(str(tmp_path), "in.bin"), os.path.join(str(tmp_path), "out.bin")
    parts = [_pitched(x).ravel(), cache.numpy().ravel(), w.numpy().ravel()] + ([b.numpy()] if bias else []) + \
            ([_pitched(sk).ravel()] if skip else [])
    np.concatenate(parts).astype(np.float32).tofile(fin)
    args = [mode, K, S, B, C, T, pre, 0.77, int(bias), int(skip), post, 0.6, fin, fout]
    r = subprocess.run([conv_exe] + [str(a) for a in args], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.returncode, r.stderr[-2000:])
    out = np.fromfile(fout, np.float32)
    Top = (T_out + 3) // 4 * 4
    y = out[:B * C * Top].reshape(B, C, Top)[:, :, :T_out]
    co = out[B * C * Top:].reshape(B, C, P)
    assert np.abs(y - ref.numpy()).max() < 1e-5
    assert np.abs(co - xin[:, :, -P:].numpy()).max() < 1e-6


@pytest.mark.parametrize("S,B,C,T,pre", [(2, 2, 3, 40, 2), (4, 1, 2, 1, 0), (5, 1, 2, 37, 1), (8, 2, 2, 600, 2),
                                          (8, 3, 150, 1, 1), (5, 2, 70, 8, 2)])   # short rows: 128 / 64 rows per CTA
def test_transposed_depthwise_kernel_source_on_cpu(conv_exe, tmp_path, S, B, C, T, pre):
    """CausalConvTranspose1d.forward causal_layers.py:183-188 (depthwise, k = 2s) with its one-sample cache."""
    g = torch.Generator().manual_seed(S * 10 + T)
    x = torch.randn(B, C, T, generator=g)
    cache = torch.randn(B, C, 1, generator=g)
    w = torch.randn(C, 1, 2 * S, generator=g)
    xin = torch.cat((cache, x if pre == 0 else F.elu(x * 0.77 if pre == 2 else x)), 2)
    ref = F.conv_transpose1d(xin.double(), w.double(), None, stride=S, padding=S, output_padding=0, groups=C)
    assert ref.shape[2] == S * T
    fin, fout = os.path.join(str(tmp_path), "in.bin"), os.path.join(str(tmp_path), "out.bin")
    np.concatenate([_pitched(x).ravel(), cache.numpy().ravel(), w.numpy().ravel()]).astype(np.float32).tofile(fin)
    args = [3, 2 * S, S, B, C, T, pre, 0.77, 0, 0, 0, 1.0, fin, fout]
    r = subprocess.run([conv_exe] + [str(a) for a in args], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.returncode, r.stderr[-2000:])
    out = np.fromfile(fout, np.float32)
    Top = (S * T + 3) // 4 * 4
    y = out[:B * C * Top].reshape(B, C, Top)[:, :, :S * T]
    co = out[B * C * Top:].reshape(B, C, 1)
    assert np.abs(y - ref.numpy()).max() < 1e-5
    assert np.abs(co - xin[:, :, -1:].numpy()).max() < 1e-6


# ------------------------------------------------------------------------------------------ gemm.cu
@pytest.fixture(scope="module")
def gemm_exe(tmp_path_factory):
    src = open(os.path.join(CSRC, "gemm.cu")).read()
    a = src.index("enum { LD_PLAIN")
    k = src.index("template <int TM, int LD, int EPI>\n__global__")
    return _build(str(tmp_path_factory.mktemp("emu_gemm")), "gemm", "gemm_extracted.inc",
                  _common_bits() + "\n" + src[a:k] + _function(src, k), "harness_gemm.cpp")


def _tm(M):
    """choose_tm() of gemm.cu."""
    if M % 128 == 0:
        return 8
    if M % 96 == 0:
        return 6
    if M % 64 == 0:
        return 4
    if M > 256:
        return 8
    return 8 if M > 96 else (6 if M > 64 else 4)


def _run_gemm(exe, tmp, LD, EPI, TM, A, Mp, M, K, B, T, pre, X, bias, R, hop, x_bs, x_rs, y_bs, y_rs, M_out):
    fin, fout = os.path.join(tmp, "in.bin"), os.path.join(tmp, "out.bin")
    parts = [A.ravel(), X.ravel()] + ([bias.ravel()] if bias is not None else []) + ([R.ravel()] if R is not None else [])
    np.concatenate(parts).astype(np.float32).tofile(fin)
    args = [LD, EPI, TM, Mp, M, K, B, T, pre, 0.8660254, int(bias is not None), int(R is not None), hop, x_bs, x_rs, y_bs,
            y_rs, M_out, fin, fout]
    r = subprocess.run([exe] + [str(v) for v in args], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.returncode, r.stderr[-2000:])
    return np.fromfile(fout, np.float32)


@pytest.mark.parametrize("B,M,K,T,pre,bias,res", [
    (2, 64, 33, 75, 0, True, True),      # TM = 4, odd K, ragged T (scalar loads), bias + residual
    (1, 96, 192, 8, 2, True, False),     # TM = 6, one hop of one stream: 8 of 128 columns used
    (3, 128, 100, 44, 1, False, False),  # TM = 8, N = 132 -> two column tiles over three streams (vector loads)
    (1, 100, 70, 5, 1, True, True),      # M not a multiple of the row tile
])
def test_ffma_gemm_source_on_cpu(gemm_exe, tmp_path, B, M, K, T, pre, bias, res):
    g = torch.Generator().manual_seed(M + K + T)
    x = torch.randn(B, K, T, generator=g)
    w = torch.randn(M, K, generator=g) / K ** 0.5
    b = torch.randn(M, generator=g) if bias else None
    r = torch.randn(B, M, T, generator=g) if res else None
    xin = x if pre == 0 else F.elu(x * 0.8660254 if pre == 2 else x)
    ref = F.conv1d(xin.double(), w.double()[:, :, None], b.double() if bias else None)
    if res:
        ref = ref + r.double()
    TM = _tm(M)
    Mp = (M + 16 * TM - 1) // (16 * TM) * (16 * TM)
    y = _run_gemm(gemm_exe, str(tmp_path), 0, 0, TM, _pack_kmajor(w.numpy(), Mp), Mp, M, K, B, T, pre, x.numpy(),
                  b.numpy() if bias else None, r.numpy() if res else None, 0, K * T, T, M * T, T, M)
    assert np.abs(y.reshape(B, M, T) - ref.numpy()).max() < 2e-5


def test_ffma_gemm_channel_last_and_stft_source_on_cpu(gemm_exe, tmp_path):
    from hilcodec_b200.weights import dft_basis
    g = torch.Generator().manual_seed(11)
    B, Fr, K, M = 2, 3, 128, 192
    q = torch.randn(B, Fr, K, generator=g)
    w = torch.randn(M, K, generator=g) / K ** 0.5
    ref = F.conv1d(q.transpose(1, 2).double(), w.double()[:, :, None])
    y = _run_gemm(gemm_exe, str(tmp_path), 1, 0, 6, _pack_kmajor(w.numpy(), M), M, M, K, B, Fr, 0, q.numpy(), None, None,
                  0, 0, 0, M * Fr, Fr, M)
    assert np.abs(y.reshape(B, M, Fr) - ref.numpy()).max() < 2e-5
    n_fft, hop, T = 64, 8, 5
    L = (T - 1) * hop + n_fft
    wav = 0.1 * torch.randn(B, 1, L, generator=g)
    basis = torch.from_numpy(dft_basis(n_fft))
    F2 = n_fft // 2 + 1
    s = F.conv1d(wav.double(), basis.double(), None, stride=hop).view(B, 2, F2, T)
    ref = s.square().sum(1).sqrt().clamp_min(1e-5).log()
    inter = torch.stack((basis[:F2, 0], basis[F2:, 0]), 1).reshape(2 * F2, n_fft)
    Mp = (2 * F2 + 95) // 96 * 96
    y = _run_gemm(gemm_exe, str(tmp_path), 2, 1, 6, _pack_kmajor(inter.numpy(), Mp), Mp, 2 * F2, n_fft, B, T, 0, wav.numpy(),
                  None, None, hop, L, 0, F2 * T, T, F2)
    err = np.abs(y.reshape(B, F2, T) - ref.numpy())
    assert np.median(err) < 1e-5 and err.max() < 1e-2


# ------------------------------------------------------------------------------------------ glue kernels
@pytest.fixture(scope="module")
def misc_exe(tmp_path_factory):
    conv = open(os.path.join(CSRC, "conv.cu")).read()
    bitp = open(os.path.join(CSRC, "bitpack.cu")).read()
    parts = [_function(conv, conv.index(sig)) for sig in (
        "__global__ void wavcat_kernel(", "template <int K>\n__global__ void conv_pre_kernel(",
        "template <int K>\n__global__ void conv_post_tanh_kernel(", "__global__ void l2norm_chlast_kernel(",
        "__global__ void chlast_to_ncw_kernel(")]
    parts += [_function(bitp, bitp.index(sig)) for sig in ("__global__ void pack_indices_kernel(",
                                                           "__global__ void unpack_indices_kernel(")]
    rvq = open(os.path.join(CSRC, "rvq.cu")).read()
    parts.append(_function(rvq, rvq.index("__global__ void kmajor_to_rows_kernel(")))
    text = "\n".join(parts)
    assert text.count("extern __shared__ float sw[];") == 2
    text = text.replace("extern __shared__ float sw[];", "float* sw = g_dyn_smem;")
    act = ("inline float apply_act_ex2(float x, int mode, float s) {\n"   # ex2.approx stand-in, differs by <= 2.4e-7
           "    if (mode == PRE_NONE) return x;\n    if (mode == PRE_SCALE_ELU) x = x * s;\n"
           "    return x > 0.f ? x : expm1f(x);\n}\n")
    return _build(str(tmp_path_factory.mktemp("emu_misc")), "misc", "misc_extracted.inc", _common_bits() + "\n" + act + text,
                  "harness_misc.cpp")


def _misc(exe, tmp, args, arrays):
    fin, fout = os.path.join(tmp, "in.bin"), os.path.join(tmp, "out.bin")
    with open(fin, "wb") as f:
        for a in arrays:
            f.write(np.ascontiguousarray(a).tobytes())
    r = subprocess.run([exe] + [str(a) for a in args] + [fin, fout], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.returncode, r.stderr[-2000:])
    return np.fromfile(fout, np.uint8)


@pytest.mark.parametrize("B,C,F", [(2, 128, 75), (1, 128, 33), (3, 40, 5)])
def test_layout_transposes_source_on_cpu(misc_exe, tmp_path, B, C, F):
    """chlast_to_ncw (latents [B, F, C] -> NCW rows, the decoder input / the tensor-core RVQ's k-major residuals) and
    kmajor_to_rows (its inverse for one block): 32 x 32 tiles through shared memory, two-dimensional blocks."""
    g = torch.Generator().manual_seed(B * 100 + F)
    q = torch.randn(B, F, C, generator=g)
    raw = _misc(misc_exe, str(tmp_path), ["transp", B, C, F], [q.numpy()]).view(np.float32)
    Fp = (F + 3) // 4 * 4
    y = raw[:B * C * Fp].reshape(B, C, Fp)
    assert np.array_equal(y[:, :, :F], q.transpose(1, 2).numpy())
    assert np.array_equal(raw[B * C * Fp:].reshape(F, C), q[0].numpy())


@pytest.mark.parametrize("n,frames", [(8, 300), (12, 129), (1, 5), (3, 77)])
def test_bitpack_kernels_source_on_cpu(misc_exe, tmp_path, n, frames):
    """Byte-exact against the numpy statement of the format, and unpack(pack(x)) == x."""
    from oracle import bitstream_oracle
    rng = np.random.default_rng(n)
    idx = rng.integers(0, 1024, size=(n, 1, frames)).astype(np.int64)
    idx[:, 0, 0] = 1023
    raw = _misc(misc_exe, str(tmp_path), ["pack", frames, n, 10], [idx])
    bpf = bitstream_oracle.bytes_per_frame(n)
    packed = raw[:frames * bpf].reshape(1, frames, bpf)
    back = raw[frames * bpf:].view(np.int64).reshape(n, 1, frames)
    assert np.array_equal(packed, bitstream_oracle.pack_numpy(idx))
    assert np.array_equal(back, idx)


def test_encoder_head_and_tail_kernels_source_on_cpu(misc_exe, tmp_path):
    g = torch.Generator().manual_seed(5)
    # wav history concat (Encoder.forward streaming.py:485-488)
    B, T, P = 2, 37, 15
    x, cache = torch.randn(B, T, generator=g), torch.randn(B, P, generator=g)
    raw = _misc(misc_exe, str(tmp_path), ["wavcat", B, T, P], [x.numpy(), cache.numpy()]).view(np.float32)
    Wp = (P + T + 3) // 4 * 4
    cat = torch.cat((cache, x), 1).numpy()
    assert np.array_equal(raw[:B * Wp].reshape(B, Wp)[:, :P + T], cat)
    assert np.array_equal(raw[B * Wp:].reshape(B, P), cat[:, -P:])
    # conv_pre (streaming.py:490): 1 -> C channels, 5 taps, on [4 history samples | chunk]
    B, C, T = 2, 6, 1030
    win = torch.randn(B, 1, T + 4, generator=g)
    w, b = torch.randn(C, 1, 5, generator=g), torch.randn(C, generator=g)
    ref = F.conv1d(win.double(), w.double(), b.double())
    raw = _misc(misc_exe, str(tmp_path), ["convpre", B, C, T], [win.numpy(), w.numpy(), b.numpy()]).view(np.float32)
    Tp = (T + 3) // 4 * 4
    assert np.abs(raw.reshape(B, C, Tp)[:, :, :T] - ref.numpy()).max() < 1e-5
    # L2Norm + transpose (streaming.py:284-285, :517)
    B, C, Fr = 3, 128, 7
    h = torch.randn(B, C, Fr, generator=g)
    h[1, :, 2] = 0.0   # zero vector: the 1e-12 clamp
    ref = (F.normalize(h.double(), p=2.0, dim=1, eps=1e-12) * 128 ** 0.5).transpose(1, 2)
    raw = _misc(misc_exe, str(tmp_path), ["l2norm", B, C, Fr, 128 ** 0.5], [_pitched(h)]).view(np.float32)
    assert np.abs(raw.reshape(B, Fr, C) - ref.numpy()).max() < 1e-5


@pytest.mark.parametrize("B,C,T,pre", [(2, 96, 45, 2), (1, 8, 1, 1), (1, 16, 2100, 0)])
def test_decoder_tail_kernel_source_on_cpu(misc_exe, tmp_path, B, C, T, pre):
    """Decoder.forward streaming.py:644-647: (scale ->) ELU -> CausalConv1d(C -> 1, k5) -> Tanh, with its cache."""
    g = torch.Generator().manual_seed(C + T)
    x = torch.randn(B, C, T, generator=g)
    cache = torch.randn(B, C, 4, generator=g)
    w = torch.randn(1, C, 5, generator=g) / (5 * C) ** 0.5
    b = torch.randn(1, generator=g)
    act = x if pre == 0 else F.elu(x * 0.70710677 if pre == 2 else x)
    xin = torch.cat((cache, act), 2)
    ref = torch.tanh(F.conv1d(xin.double(), w.double(), b.double()))
    raw = _misc(misc_exe, str(tmp_path), ["convpost", B, C, T, pre, 0.70710677],
                [_pitched(x), cache.numpy(), w.numpy(), b.numpy()]).view(np.float32)
    assert np.abs(raw[:B * T].reshape(B, 1, T) - ref.numpy()).max() < 1e-5
    assert np.abs(raw[B * T:].reshape(B, C, 4) - xin[:, :, -4:].numpy()).max() < 1e-6
