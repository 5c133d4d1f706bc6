"""Per-kernel parity (-m gpu): each CUDA operator through the C ABI against the torch CPU
op the reference calls (F.conv1d / F.conv_transpose1d / elu / log ...).  fp32 tolerance is
written per test; cache outputs are pure copies and must be bit-exact."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from hilcodec_b200 import _lib

pytestmark = pytest.mark.gpu


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _pre_ref(x, pre, s):
    if pre == 0:
        return x
    if pre == 2:
        x = x * s
    return F.elu(x)


@pytest.mark.parametrize("B,Cc,T,K,S,pre,bias,skip", [
    (2, 64, 256, 5, 1, 0, True, False),
    (3, 96, 75, 5, 1, 1, True, True),     # T not a multiple of 4 -> scalar path
    (2, 128, 1, 5, 1, 0, True, False),    # one frame per call: cache shift-mix
    (1, 32, 2, 5, 1, 2, False, False),
    (2, 128, 320, 4, 2, 0, True, False),
    (2, 256, 160, 8, 4, 0, True, False),
    (2, 64, 40, 10, 5, 0, True, False),
    (2, 64, 8, 16, 8, 0, True, False),
    (2, 512, 600, 10, 5, 0, True, False),   # vectorised strided kernel, stride 5 (unaligned history)
    (1, 1024, 600, 16, 8, 0, True, False),  # 75 outputs: ragged last group of 4
    (2, 128, 1000, 4, 2, 1, False, True),   # activation + skip through the vectorised kernel
    (1, 256, 72, 8, 4, 0, True, False),     # 18 outputs
    (1, 1024, 4, 5, 1, 1, False, False),
    (2, 192, 1000, 5, 1, 0, True, True),
])
def test_dwconv(B, Cc, T, K, S, pre, bias, skip):
    lib = _lib.load()
    g = torch.Generator().manual_seed(B * 1000 + T)
    P = K - S
    x = torch.randn(B, Cc, T, generator=g)
    cache = torch.randn(B, Cc, P, generator=g)
    w = torch.randn(Cc, 1, K, generator=g) / K ** 0.5
    b = torch.randn(Cc, generator=g) if bias else None
    xin = torch.cat((cache, _pre_ref(x, pre, 0.77)), 2)
    y_ref = F.conv1d(xin, w, b, stride=S, groups=Cc)
    sk = torch.randn_like(y_ref) if skip else None
    if skip:
        y_ref = y_ref + sk
    c_ref = xin[:, :, -P:]
    xd, cd, wd = x.cuda(), cache.cuda(), w.cuda()
    bd = b.cuda() if bias else None
    skd = sk.cuda() if skip else None
    y = torch.empty(y_ref.shape, device="cuda")
    co = torch.empty(B, Cc, P, device="cuda")
    _lib.check(lib.hil_op_dwconv(_ptr(xd), _ptr(cd), _ptr(co), _ptr(wd), _ptr(bd), _ptr(skd), _ptr(y),
                                 B, Cc, T, K, S, pre, 0.77, _stream()))
    torch.cuda.synchronize()
    assert torch.allclose(y.cpu(), y_ref, rtol=1e-5, atol=2e-6), (y.cpu() - y_ref).abs().max()
    assert torch.allclose(co.cpu(), c_ref, rtol=0, atol=1e-6)
    if pre == 0:
        assert torch.equal(co.cpu(), c_ref)


@pytest.mark.parametrize("B,Cc,T,S,pre", [(2, 192, 300, 2, 1), (2, 384, 75, 4, 2), (1, 768, 15, 5, 2),
                                           (3, 1536, 1, 8, 1), (2, 64, 7, 8, 0)])
def test_dwconv_transpose(B, Cc, T, S, pre):
    lib = _lib.load()
    g = torch.Generator().manual_seed(S * 100 + T)
    x = torch.randn(B, Cc, T, generator=g)
    cache = torch.randn(B, Cc, 1, generator=g)
    w = torch.randn(Cc, 1, 2 * S, generator=g)
    xin = torch.cat((cache, _pre_ref(x, pre, 0.7071)), 2)
    y_ref = F.conv_transpose1d(xin, w, None, stride=S, padding=S, output_padding=0, groups=Cc)
    assert y_ref.shape[2] == T * S
    y = torch.empty(B, Cc, T * S, device="cuda")
    co = torch.empty(B, Cc, 1, device="cuda")
    xd, cd, wd = x.cuda(), cache.cuda(), w.cuda()  # keep the device tensors alive across the call
    _lib.check(lib.hil_op_dwconv_transpose(_ptr(xd), _ptr(cd), _ptr(co), _ptr(wd), _ptr(y),
                                           B, Cc, T, S, pre, 0.7071, _stream()))
    torch.cuda.synchronize()
    assert torch.allclose(y.cpu(), y_ref, rtol=1e-5, atol=2e-6), (y.cpu() - y_ref).abs().max()
    assert torch.allclose(co.cpu(), xin[:, :, -1:], rtol=0, atol=1e-6)


@pytest.mark.parametrize("B,M,K,T,pre,bias,res", [
    (2, 64, 64, 512, 1, False, False),
    (2, 128, 64, 300, 2, False, False),
    (3, 96, 192, 75, 0, True, False),      # ragged T, flattened columns
    (2, 64, 33, 640, 0, True, True),       # SpecBlock 1x1: odd K, bias, residual add
    (1, 1024, 513, 75, 0, True, True),
    (2, 768, 1536, 40, 0, True, False),
    (1, 128, 1024, 5, 0, True, False),
    (4, 192, 192, 1, 1, False, False),     # one column per stream
    (1, 256, 257, 129, 2, True, True),
    # one stream, one hop (the streaming shapes: gemm_skinny.cu)
    (1, 768, 1536, 8, 2, True, True),
    (1, 1536, 128, 1, 0, False, False),
    (3, 384, 768, 40, 1, False, True),
    (64, 96, 96, 5, 1, True, False),
    # flat tiles (128 / T whole clips per tile) for short chunks of many streams
    (64, 1024, 512, 8, 0, True, False),
    (64, 768, 1536, 8, 2, True, True),
    (50, 256, 257, 8, 1, False, True),    # odd K, clips not a multiple of 16
    (40, 192, 384, 4, 1, True, False),    # 4 samples per clip: 32 clips per tile
    (9, 128, 96, 16, 0, False, True),     # 16 samples per clip, 144 columns: second tile nearly empty
    # 64 concurrent streams, one hop: 40-column chunks take the tensor-core tiles (one partly filled tile per clip)
    (64, 256, 256, 40, 1, False, False),
    (64, 384, 768, 40, 0, True, True),
    (64, 512, 512, 8, 1, False, False),    # 512 columns: the skinny kernel's largest launch
])
def test_pointwise(B, M, K, T, pre, bias, res):
    lib = _lib.load()
    g = torch.Generator().manual_seed(M + K + T)
    x = torch.randn(B, K, T, generator=g)
    w = (torch.randn(M, K, 1, generator=g) / K ** 0.5).contiguous()
    b = torch.randn(M, generator=g) if bias else None
    r = torch.randn(B, M, T, generator=g) if res else None
    y_ref = F.conv1d(_pre_ref(x, pre, 0.8660254), w, b)
    y64 = F.conv1d(_pre_ref(x, pre, 0.8660254).double(), w.double(), b.double() if bias else None)
    if res:
        y_ref = y_ref + r
        y64 = y64 + r.double()
    y = torch.empty(B, M, T, device="cuda")
    xd = x.cuda()
    bd = b.cuda() if bias else None
    rd = r.cuda() if res else None
    _lib.check(lib.hil_op_pointwise(_ptr(xd), _ptr(w), _ptr(bd), _ptr(rd), _ptr(y), B, M, K, T, pre, 0.8660254,
                                    _stream()))
    torch.cuda.synchronize()
    err = (y.cpu().double() - y64).abs().max().item()
    ref_err = (y_ref.double() - y64).abs().max().item()
    # As accurate as the CPU fp32 result is (both are fp32 accumulations in different orders), plus what the tensor
    # pipe's accumulator costs: tcgen05.mma adds each k16 step into the fp32 TMEM accumulator with TRUNCATION, a bias
    # that grows linearly with K (measured on the B200: 4.8e-6 at K = 768, 8.5e-6 at K = 1536 for O(1) outputs, i.e.
    # ~K * 6e-9 typical, 9e-9 worst seen; the FP32 kernels that take the short chunks stay at the CPU's level).
    # DESIGN.md section 5.
    assert err <= max(4 * ref_err, 2e-6, 1.2e-8 * K), (err, ref_err)


@pytest.mark.parametrize("B,n_fft,hop,T", [(2, 64, 1, 640), (2, 128, 2, 320), (1, 256, 8, 75), (64, 256, 8, 40),
                                            (64, 512, 40, 8), (37, 512, 40, 8), (48, 1024, 320, 4),   # flat tiles: whole clips per tile
                                            (2, 512, 40, 15), (2, 1024, 320, 3), (1, 1024, 320, 1),
                                            # tensor-core kernel (T >= 64, 16-byte aligned rows)
                                            (2, 64, 1, 1000), (2, 128, 2, 500), (2, 256, 8, 300),
                                            (2, 512, 40, 132), (3, 1024, 320, 76), (1, 1024, 320, 64)])
def test_stft_logmag(B, n_fft, hop, T):
    from hilcodec_b200.weights import dft_basis
    lib = _lib.load()
    g = torch.Generator().manual_seed(n_fft + T)
    L = (T - 1) * hop + n_fft
    wav = 0.1 * torch.randn(B, 1, L, generator=g)
    wav[0, 0, : L // 3] = 0.0  # silence: exercises the 1e-5 clamp
    w = torch.from_numpy(dft_basis(n_fft)).contiguous()
    s = F.conv1d(wav, w, None, stride=hop)
    Fr = n_fft // 2 + 1
    s = s.view(B, 2, Fr, T)
    y_ref = s.square().sum(1).sqrt().clamp_min(1e-5).log()
    s64 = F.conv1d(wav.double(), w.double(), None, stride=hop).view(B, 2, Fr, T)
    y64 = s64.square().sum(1).sqrt().clamp_min(1e-5).log()
    y = torch.empty(B, Fr, T, device="cuda")
    wd = wav.cuda()
    _lib.check(lib.hil_op_stft_logmag(_ptr(wd), _ptr(w), _ptr(y), B, n_fft, hop, T, _stream()))
    torch.cuda.synchronize()
    # log of a magnitude near cancellation amplifies fp32 noise: compare where the fp64
    # magnitude is well conditioned, and bound everything by the CPU fp32 error itself
    err = (y.cpu().double() - y64).abs()
    ref_err = (y_ref.double() - y64).abs()
    # (the tensor-core kernel's 3xTF32 DFT is ~2-3x the fp32 rounding error; the worst element sits where
    # the magnitude nearly cancels and the log amplifies it)
    assert err.max().item() <= max(16 * ref_err.max().item(), 1e-4), (err.max().item(), ref_err.max().item())
    assert torch.median(err).item() < 5e-6  # 3xTF32 on the tensor-core path: ~2e-6 on the log-magnitude


@pytest.mark.parametrize("B,Cc,T,skip,pre", [
    (2, 96, 2400, True, 1),     # fused tensor-core kernel, in-place residual (TMA reduce-add)
    (3, 64, 1000, False, 2),    # ragged last tile, pre-scale + ELU prologue
    (1, 192, 248, False, 0),    # exactly two 124-column tiles
    (2, 384, 130, True, 1),     # multi row-tile, second tile nearly empty
    (2, 128, 75, False, 1),     # T not a multiple of 4 -> unfused FFMA + depthwise fallback
    (4, 96, 8, True, 1),        # short chunk (streaming): depthwise in the skinny GEMM's epilogue
    (64, 256, 40, True, 1),     # 64 streams, one hop: fused tensor-core kernel on a 40-column chunk (44 of 128 tile columns)
    (64, 384, 40, False, 2),
    (33, 128, 36, True, 1),     # shortest chunk the tensor-core tiles take (T >= 32, B * T > 512)
    # flat tiles: 16 whole 8-sample clips per 128-column tile ({t, clip, k} tensor map), every clip's window from its cache
    (64, 768, 8, True, 0),      # decoder stage 1 at 64 streams: unit 1 of a ResBlock (in-place residual)
    (64, 512, 8, False, 2),     # encoder stage 3: unit 0 (pre-scale + ELU prologue)
    (37, 192, 8, True, 1),      # clips not a multiple of 4 / 16: partly filled last chunk and tile
    (16, 256, 8, False, 1),     # exactly one tile
])
def test_dws_block(B, Cc, T, skip, pre):
    """DWSBlock (ELU -> 1x1 -> depthwise k5 + bias) plus the ResBlock's residual add."""
    lib = _lib.load()
    g = torch.Generator().manual_seed(Cc + T)
    x = torch.randn(B, Cc, T, generator=g)
    w = (torch.randn(Cc, Cc, 1, generator=g) / Cc ** 0.5).contiguous()
    wd = torch.randn(Cc, 1, 5, generator=g) / 5 ** 0.5
    bd = torch.randn(Cc, generator=g)
    cache = torch.randn(B, Cc, 4, generator=g)
    sk = torch.randn(B, Cc, T, generator=g) if skip else None
    pw = F.conv1d(_pre_ref(x, pre, 0.8660254).double(), w.double())
    xin = torch.cat((cache.double(), pw), 2)
    ref = F.conv1d(xin, wd.double(), bd.double(), groups=Cc)
    if skip:
        ref = ref + sk.double()
    xd, wdd, bdd, cd = x.cuda(), wd.cuda(), bd.cuda(), cache.cuda()
    y = sk.cuda() if skip else torch.empty(B, Cc, T, device="cuda")   # residual accumulates in place, like h
    tmp = torch.empty(B, Cc, T, device="cuda")
    co = torch.empty(B, Cc, 4, device="cuda")
    _lib.check(lib.hil_op_dws(_ptr(xd), _ptr(w), _ptr(wdd), _ptr(bdd), _ptr(cd), _ptr(co), _ptr(y if skip else None),
                              _ptr(tmp), _ptr(y), B, Cc, T, pre, 0.8660254, 0, 1.0, _stream()))
    torch.cuda.synchronize()
    assert (y.cpu().double() - ref).abs().max().item() < 2e-5
    assert (co.cpu().double() - xin[:, :, -4:]).abs().max().item() < 1e-5


def _resblock_ref(x, w0, w1, d0w, d0b, d1w, d1b, c0, c1, pre, pre_scale):
    """ResBlock.forward streaming.py:252-275 (merged scaling) in fp64, with both depthwise caches."""
    C = x.shape[1]
    a = _pre_ref(x, pre, pre_scale).double()
    p0 = torch.cat((c0.double(), F.conv1d(a, w0.double())), 2)
    a = F.elu(F.conv1d(p0, d0w.double(), d0b.double(), groups=C))
    p1 = torch.cat((c1.double(), F.conv1d(a, w1.double())), 2)
    y = x.double() + F.conv1d(p1, d1w.double(), d1b.double(), groups=C)
    return y, p0[:, :, -4:], p1[:, :, -4:]


@pytest.mark.parametrize("B,Cc,T,pre", [
    (2, 96, 2400, 1),     # decoder stage 3 shape: one m-block, 128-column tiles (120 outputs each), 20 tiles
    (1, 64, 1000, 2),     # encoder stage 0: half-empty m-block, ragged last tile, scaled ELU prologue
    (3, 128, 128, 1),     # shortest chunk the fused kernel takes: two tiles, the second one 8 columns wide
    (2, 128, 364, 1),     # three full tiles and a 4-column one
    (1, 32, 200, 1),
    (4, 64, 500, 2),      # even batch at C = 64: the clip-pair form (block-diagonal weights, see hil_op_resblock)
    (2, 192, 500, 1),     # 128 < C <= 256: two row blocks per 64-column tile (HILCODEC_RB_WIDE=1 only)
    (1, 256, 300, 2),
])
def test_resblock_fused(B, Cc, T, pre):
    """gemm_rb.cu (whole ResBlock in one kernel, h updated in place) against the fp64 reference and against the two
    fused-DWS launches it replaces: same arithmetic, so the two CUDA paths must agree bit for bit."""
    import os
    if Cc > 128 and os.environ.get("HILCODEC_RB_WIDE") != "1":
        pytest.skip("the 64-column ResBlock variant is opt-in (HILCODEC_RB_WIDE=1)")
    lib = _lib.load()
    g = torch.Generator().manual_seed(Cc * 31 + T)
    x = torch.randn(B, Cc, T, generator=g)
    w0 = (torch.randn(Cc, Cc, 1, generator=g) / Cc ** 0.5).contiguous()
    w1 = (torch.randn(Cc, Cc, 1, generator=g) / Cc ** 0.5).contiguous()
    d0w = torch.randn(Cc, 1, 5, generator=g) / 5 ** 0.5
    d1w = 0.57735 * torch.randn(Cc, 1, 5, generator=g) / 5 ** 0.5
    d0b, d1b = torch.randn(Cc, generator=g), torch.randn(Cc, generator=g)
    c0, c1 = torch.randn(B, Cc, 4, generator=g), torch.randn(B, Cc, 4, generator=g)
    y_ref, c0_ref, c1_ref = _resblock_ref(x, w0, w1, d0w, d0b, d1w, d1b, c0, c1, pre, 0.8660254)
    dev = [t.cuda() for t in (d0w, d0b, d1w, d1b, c0, c1)]
    outs = []
    for fused in (1, 0):
        h = x.cuda().clone()
        c0o, c1o = torch.zeros(B, Cc, 4, device="cuda"), torch.zeros(B, Cc, 4, device="cuda")
        t1, t2 = torch.zeros(B, Cc, T, device="cuda"), torch.zeros(B, Cc, T, device="cuda")
        _lib.check(lib.hil_op_resblock(_ptr(h), _ptr(w0), _ptr(w1), _ptr(dev[0]), _ptr(dev[1]), _ptr(dev[2]), _ptr(dev[3]),
                                       _ptr(dev[4]), _ptr(c0o), _ptr(dev[5]), _ptr(c1o), _ptr(t1), _ptr(t2),
                                       B, Cc, T, pre, 0.8660254, fused, _stream()))
        torch.cuda.synchronize()
        outs.append((h.cpu(), c0o.cpu(), c1o.cpu()))
    for h, c0o, c1o in outs:
        assert (h.double() - y_ref).abs().max().item() < 3e-5
        assert (c0o.double() - c0_ref).abs().max().item() < 1e-5
        assert (c1o.double() - c1_ref).abs().max().item() < 1e-5
    assert torch.equal(outs[0][0], outs[1][0]), (outs[0][0] - outs[1][0]).abs().max()
    assert torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][2], outs[1][2])


@pytest.mark.parametrize("B,K,M,T_in,S,pre", [
    (2, 192, 96, 300, 2, 2),     # decoder stage 3: 192 -> 96, x2, scaled-ELU prologue
    (1, 384, 192, 200, 4, 2),    # stage 2: two row tiles
    (2, 768, 384, 60, 5, 2),     # stage 1: stride 5 (tiles do not start on an input boundary)
    (1, 1536, 768, 40, 8, 0),    # stage 0: no activation (applied by the producer), six row tiles
    (3, 64, 32, 68, 2, 2),       # ragged: 136 output columns, second tile 8 wide
])
def test_upsample_fused(B, K, M, T_in, S, pre):
    """Decoder upsampling layer (act -> CausalConvTranspose1d depthwise -> 1x1 + bias) as one tensor-core kernel,
    against the fp64 reference and the two-kernel path."""
    lib = _lib.load()
    g = torch.Generator().manual_seed(K + 7 * S + T_in)
    x = torch.randn(B, K, T_in, generator=g)
    if (T_in % 4) != 0:
        pytest.skip("dense-row operator entry needs T_in % 4 == 0")
    cache = torch.randn(B, K, 1, generator=g)
    wu = torch.randn(K, 1, 2 * S, generator=g) / (2 * S) ** 0.5
    wp = (torch.randn(M, K, 1, generator=g) / K ** 0.5).contiguous()
    bias = torch.randn(M, generator=g)
    xin = torch.cat((cache, _pre_ref(x, pre, 0.7071)), 2).double()
    u = F.conv_transpose1d(xin, wu.double(), None, stride=S, padding=S, output_padding=0, groups=K)
    y_ref = F.conv1d(u, wp.double(), bias.double())
    T = S * T_in
    assert y_ref.shape[2] == T
    xd, cd, wud, bd = x.cuda(), cache.cuda(), wu.cuda(), bias.cuda()
    outs = []
    for fused in (1, 0):   # one kernel / transposed conv + 1x1 through the fp32 intermediate
        y = torch.zeros(B, M, T, device="cuda")
        co = torch.zeros(B, K, 1, device="cuda")
        tmp = torch.zeros(B, K, T, device="cuda")
        _lib.check(lib.hil_op_upsample(_ptr(xd), _ptr(cd), _ptr(co), _ptr(wud), _ptr(wp), _ptr(bd), _ptr(tmp), _ptr(y),
                                       B, K, M, T_in, S, pre, 0.7071, fused, _stream()))
        torch.cuda.synchronize()
        outs.append((y.cpu(), co.cpu()))
    scale = max(1.0, y_ref.abs().max().item())
    for y, co in outs:
        assert (y.double() - y_ref).abs().max().item() < 2e-5 * scale
        assert (co.double() - xin[:, :, -1:]).abs().max().item() < 1e-6
    assert (outs[0][0] - outs[1][0]).abs().max().item() < 1e-5 * scale


@pytest.mark.parametrize("B,K,M,T,r,pre", [
    (2, 64, 128, 2400, 2, 2),     # encoder stage 0: 64 -> 128, stride 2 (20 tiles of 120 columns per clip)
    (1, 128, 256, 1200, 4, 2),    # stage 1: stride 4, two row tiles
    (2, 256, 512, 600, 5, 2),     # stage 2: stride 5 (tile = 8 halo + 120 new columns), four row tiles
    (3, 64, 128, 136, 2, 1),      # ragged: second tile 16 columns wide
    (1, 32, 64, 240, 4, 0),       # half-empty row tile, no activation, last tile partly past the end
    (2, 96, 192, 200, 5, 2),
])
def test_downsample_fused(B, K, M, T, r, pre):
    """Encoder downsampling pair (act -> 1x1 -> causal strided depthwise conv + bias) as one tensor-core kernel, against
    the fp64 reference and the two-kernel path (same FMA order in the depthwise window: the two CUDA paths agree to the
    last bits of the 1x1 result, which the same GEMM mainloop produces)."""
    lib = _lib.load()
    g = torch.Generator().manual_seed(K + 13 * r + T)
    x = torch.randn(B, K, T, generator=g)
    wp = (torch.randn(M, K, 1, generator=g) / K ** 0.5).contiguous()
    wd = torch.randn(M, 1, 2 * r, generator=g) / (2 * r) ** 0.5
    bd = torch.randn(M, generator=g)
    cache = torch.randn(B, M, r, generator=g)
    pw = F.conv1d(_pre_ref(x, pre, 0.7745967).double(), wp.double())
    xin = torch.cat((cache.double(), pw), 2)
    y_ref = F.conv1d(xin, wd.double(), bd.double(), stride=r, groups=M)
    T2 = T // r
    assert y_ref.shape[2] == T2
    if T % 4 or T2 % 4:
        pytest.skip("dense-row operator entry needs T % 4 == 0 and (T / r) % 4 == 0")
    xd, cd, wdd, bdd = x.cuda(), cache.cuda(), wd.cuda(), bd.cuda()
    outs = []
    for fused in (1, 0):
        y = torch.zeros(B, M, T2, device="cuda")
        co = torch.zeros(B, M, r, device="cuda")
        tmp = torch.zeros(B, M, T, device="cuda")
        _lib.check(lib.hil_op_downsample(_ptr(xd), _ptr(cd), _ptr(co), _ptr(wp), _ptr(wdd), _ptr(bdd), _ptr(tmp), _ptr(y),
                                         B, K, M, T, r, pre, 0.7745967, fused, _stream()))
        torch.cuda.synchronize()
        outs.append((y.cpu(), co.cpu()))
    scale = max(1.0, y_ref.abs().max().item())
    for y, co in outs:
        assert (y.double() - y_ref).abs().max().item() < 2e-5 * scale
        assert (co.double() - xin[:, :, -r:]).abs().max().item() < 1e-5
    assert torch.equal(outs[0][0], outs[1][0]), (outs[0][0] - outs[1][0]).abs().max()
    assert torch.equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("size,B,F,n,train", [(1024, 4, 600, 12, False), (1024, 32, 75, 8, True), (256, 1, 2048, 3, False),
                                               (1024, 2, 17000, 2, False)])
def test_rvq_tensor_core_search_bit_identical(size, B, F, n, train):
    """Batches (>= 2048 frames) run the search as one fp32-accurate tensor-core GEMM per stage plus a decision kernel
    that re-scores near ties with the FFMA expression (rvq.cu rvq_tc_select_kernel): indices and dequantised sums
    must equal the one-kernel FFMA search bit for bit -- exact ties, realistic latents drawn near codebook rows (small
    margins in the later stages), and more than one 32 768-frame chunk."""
    from hilcodec_b200 import streaming as S, weights as W, _lib
    lib = _lib.load()
    cfg = W.CodecConfig(num_quantizers=n, codebook_size=size)
    g = torch.Generator().manual_seed(size + F)
    cbs = {f"quantizer.layers.{i}.embed": (torch.randn(size, 128, generator=g) * 0.6 ** i).numpy() for i in range(n)}
    cbs["quantizer.layers.0.embed"][7] = cbs["quantizer.layers.0.embed"][3]
    cbs["quantizer.layers.0.embed"][130] = cbs["quantizer.layers.0.embed"][3]
    core = S._NativeCodec(cfg, _lib.HIL_GRAPH_TRAIN if train else _lib.HIL_GRAPH_DEPLOY)
    core.set_weights(cbs)
    z = torch.nn.functional.normalize(torch.randn(B, F, 128, generator=g), dim=2) * 128 ** 0.5
    # a quarter of the frames are sums of codebook rows plus a little noise: the later stages see residuals that are tiny
    # against the codebook scale, where margins are smallest
    pick = torch.randint(0, size, (n, B, F // 4), generator=g)
    acc = sum(torch.from_numpy(cbs[f"quantizer.layers.{i}.embed"])[pick[i]] for i in range(n))
    z[:, : F // 4] = acc + 1e-3 * torch.randn(B, F // 4, 128, generator=g)
    z[0, 0] = torch.from_numpy(cbs["quantizer.layers.0.embed"][3])   # an exact three-way tie: the first index wins
    z = z.cuda()
    prev = lib.hil_set_tensor_cores(1 | 16)
    try:
        idx_tc, q_tc = core.rvq_encode(z, n, with_sum=True)
        only_idx = core.rvq_encode(z, n)
        lib.hil_set_tensor_cores(1 | 16 | 256)
        idx_ff, q_ff = core.rvq_encode(z, n, with_sum=True)
        torch.cuda.synchronize()
    finally:
        lib.hil_set_tensor_cores(prev)
    assert torch.equal(idx_tc, idx_ff)
    assert torch.equal(only_idx, idx_ff)
    assert torch.equal(q_tc, q_ff)
    assert idx_tc[0, 0, 0].item() == 3


@pytest.mark.parametrize("size,frames,n,train", [(1024, 1, 12, False), (1024, 64, 12, False), (1024, 75, 8, True),
                                                  (1024, 1000, 3, False), (200, 33, 2, False), (64, 5, 4, True),
                                                  (1024, 200, 4, False)])   # 1 / 2 / 4 frames per warp: <= 128 / <= 256 / more
def test_rvq_few_frame_variants_bit_identical(size, frames, n, train, monkeypatch):
    """The streaming RVQ paths (one cluster launch with DSMEM candidate exchange; n + 1 per-stage launches) against the
    one-kernel search: same indices and the same dequantised sum, bit for bit, including exact ties."""
    import subprocess, sys, json, os
    code = r"""
import json, sys, torch
sys.path.insert(0, %r)
from hilcodec_b200 import streaming as S, weights as W, _lib
size, frames, n, train = %d, %d, %d, %d
cfg = W.CodecConfig(num_quantizers=n, codebook_size=size)
g = torch.Generator().manual_seed(size + frames)
cbs = {f"quantizer.layers.{i}.embed": (torch.randn(size, 128, generator=g) * 0.7 ** i).numpy() for i in range(n)}
cbs["quantizer.layers.0.embed"][7] = cbs["quantizer.layers.0.embed"][3]
core = S._NativeCodec(cfg, _lib.HIL_GRAPH_TRAIN if train else _lib.HIL_GRAPH_DEPLOY)
core.set_weights(cbs)
z = torch.nn.functional.normalize(torch.randn(1, frames, 128, generator=g), dim=2) * 128 ** 0.5
z[0, 0] = torch.from_numpy(cbs["quantizer.layers.0.embed"][3])
idx, q = core.rvq_encode(z.cuda(), n, with_sum=True)
torch.cuda.synchronize()
print(json.dumps({"idx": idx.cpu().flatten().tolist(), "q": q.cpu().flatten().tolist()}))
""" % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), size, frames, n, int(train))
    outs = {}
    for name, env in (("cluster", {}), ("one_kernel", {"HILCODEC_RVQ_CLUSTER": "0"})):
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300,
                           env={**os.environ, **env})
        assert r.returncode == 0, r.stderr[-2000:]
        outs[name] = json.loads(r.stdout.strip().splitlines()[-1])
    assert outs["cluster"]["idx"] == outs["one_kernel"]["idx"]
    assert outs["cluster"]["q"] == outs["one_kernel"]["q"]
    assert outs["cluster"]["idx"][0] == 3          # frame 0 IS code 3 of stage 0; its exact copy (code 7) must lose
