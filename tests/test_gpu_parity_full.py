"""Full-size and edge-case parity (-m gpu) of the CUDA path against the CPU oracle, with the PUBLISHED weights.

SURVEY.md section 8d "Acceptance": indices bit-exact on BASELINE configs 2 and 3 under the near-tie policy (7.3:
a disagreement with the CPU result is accepted only where the fp64 relative gap between the two best codes is
< 1e-5, judged at the FIRST differing stage of a frame), latents and PCM max-abs-err < 1e-4.  The oracle port runs
~1 000 frames/s on the GPU box's host cores, so config 2 (4 800 frames) costs ~5 s and config 3 (19 200) ~20 s.

Edge-case inputs the reference's graph treats specially: all-zero audio (the `clamp(1e-5).log()` floor of every
SpecBlock, streaming.py:351), full-scale and clipped audio, DC, a single impulse, near-silence.
"""
import os

import numpy as np
import pytest
import torch

from hilcodec_b200 import streaming as S
from hilcodec_b200 import weights as W
from oracle import hilcodec_oracle as O

from helpers import index_report, oracle_cfg, params, synth_wav

pytestmark = pytest.mark.gpu

TOL = 1e-4          # BASELINE.json north_star: latents / PCM max-abs-err
GAP = 1e-5          # near-tie policy


def _weights(name):
    cfg = W.CONFIGS[name]
    if W.have_pretrained(name):
        return cfg, W.load_pretrained(name), True
    return cfg, W.random_weights(cfg, 4), False


def _oracle(cfg, p, x, n_q, sub=16):
    """The oracle in sub-batches (a 256-clip batch would hold ~10 GB of fp32 activations on the host)."""
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    outs = []
    with torch.no_grad():
        for b in range(0, x.shape[0], sub):
            outs.append(O.codec_forward(oracle_cfg(n_q), p, x[b:b + sub], n_q))
    return {"z": torch.cat([o["z"] for o in outs]), "indices": torch.cat([o["indices"] for o in outs], 1),
            "q": torch.cat([o["q"] for o in outs]), "wav": torch.cat([o["wav"] for o in outs])}


def _check_against_oracle(m, p, x, n_q, o, what):
    """Bars: latents / PCM max-abs-err < 1e-4 against the reference's fp32 arithmetic (the oracle), indices equal under
    the near-tie policy.  Conditioning: a SpecBlock takes log(max(|STFT|, 1e-5)) (streaming.py:351); on inputs whose
    spectrum has exactly-empty bins (DC, a bin-centred sine) those bins hold only fp32 ROUNDING NOISE, which for the
    long windows lies above the 1e-5 floor, and its logarithm is O(1) and implementation-defined -- the reference's own
    fp32 result is then 1e-2 away from the fp64 evaluation of the same graph and disagrees with it on 10-35 % of the
    indices (measured: tests/test_gpu_parity_full.py history).  So a clip whose latents differ from
    the fp32 oracle by >= 1e-4 is re-evaluated in fp64: the CUDA result must be no farther from the fp64 truth than 4x
    the reference's own fp32 result is (or 1e-4), and for such a clip the indices are judged against the search run on
    the CUDA path's OWN latents (a codebook decision cannot be compared across different latents)."""
    ocfg = oracle_cfg(n_q)
    xd = x.cuda()
    idx, y = m.codec_forward(xd, n_q)                       # the fused C-ABI call the bench times
    ce, cd = m.initialize_cache(xd)
    z, _ = m.encoder(xd, *ce)
    idx4 = m.quantizer(z, n_q)
    assert torch.equal(idx4, idx), what                     # four-call flow == fused call
    assert torch.isfinite(y).all() and torch.isfinite(z).all(), what
    zc, idc = z.cpu(), idx.cpu()
    z_clip = (zc - o["z"]).abs().amax(dim=(1, 2))
    ill = torch.nonzero(z_clip >= TOL).flatten().tolist()
    cond = {}
    if ill:
        # fp64 evaluation of the ENCODER for every clip of the batch: the reference's own fp32 noise has a heavy tail
        # (hil_music, 256 noise clips: median 3e-6, but single clips at 6e-5 .. 1e-4 where one near-empty STFT bin
        # meets the log), so a clip is judged against its own conditioning AND against that tail
        p64 = O.to_dtype(p, torch.float64)
        e_cpu = torch.zeros(x.shape[0], dtype=torch.float64)
        z64s = {}
        with torch.no_grad():
            for b0 in range(0, x.shape[0], 8):
                z64 = O.encoder_forward(ocfg, p64, x[b0:b0 + 8].double(), None)[0]
                e_cpu[b0:b0 + 8] = (o["z"][b0:b0 + 8].double() - z64).abs().amax(dim=(1, 2))
                for b in ill:
                    if b0 <= b < b0 + 8:
                        z64s[b] = z64[b - b0]
        tail = 4.0 * e_cpu.max().item()
        for b in ill:
            e_gpu = (zc[b].double() - z64s[b]).abs().max().item()
            cond[b] = (e_gpu, e_cpu[b].item())
            assert e_gpu <= max(TOL, 8.0 * e_cpu[b].item(), tail), \
                (what, "clip", b, "cuda vs fp64", e_gpu, "reference fp32 vs fp64", e_cpu[b].item(), "batch tail x4", tail)
        cond["reference_fp32_vs_fp64_over_batch"] = (f"max {e_cpu.max().item():.2e}", f"median {e_cpu.median().item():.2e}")
    good = torch.ones(x.shape[0], dtype=torch.bool)
    good[ill] = False
    z_err = z_clip[good].max().item() if good.any() else 0.0
    # indices: near-tie policy against the oracle on well-conditioned clips ...
    bad = worst = 0
    if good.any():
        gi = torch.nonzero(good).flatten()
        bad, worst = index_report(ocfg, p, z[gi.cuda()], idx[:, gi.cuda()], o["indices"][:, gi], n_q)
        assert bad == 0 or worst < GAP, (what, bad, worst)
    # ... and against the search on the CUDA path's own latents on the others
    if ill:
        ii = torch.tensor(ill)
        with torch.no_grad():
            own = O.rvq_encode(ocfg, p, zc[ii], n_q)
        b2, w2 = index_report(ocfg, p, z[ii.cuda()], idx[:, ii.cuda()], own, n_q)
        assert b2 == 0 or w2 < GAP, (what, "ill-conditioned clips", ill, b2, w2)
    # PCM: clips whose indices all agree must decode to the oracle's PCM; the decoder alone, fed the oracle's own
    # dequantised latents, must do so for EVERY clip (isolates decoder parity from the encoder side)
    clean = ~(idc != o["indices"]).any(dim=0).any(dim=1)
    y_err = (y.cpu() - o["wav"])[clean].abs().max().item() if clean.any() else 0.0
    assert y_err < TOL, (what, y_err)
    y2, _ = m.decoder(o["q"].cuda(), *cd)
    d_err = (y2.cpu() - o["wav"]).abs().max().item()
    assert d_err < TOL, (what, d_err)
    return {"frames": idx.shape[1] * idx.shape[2], "near_tie_frames": bad, "worst_gap": worst, "z_err": z_err,
            "pcm_err": y_err, "decoder_err": d_err, "clean_clips": int(clean.sum()),
            "ill_conditioned_clips": {b: (f"cuda-fp64 {e[0]:.2e}", f"ref32-fp64 {e[1]:.2e}") if isinstance(b, int) else e
                                      for b, e in cond.items()}}


def test_config2_full_size_vs_oracle():
    """BASELINE configs[1]: hil_speech, 64 x 24000, n_q = 8 -- every one of the 38 400 decisions against the oracle."""
    cfg, w, _ = _weights("hil_speech")
    p = params(w)
    x = synth_wav(64, 24000, seed=1234)
    o = _oracle(cfg, p, x, 8)
    r = _check_against_oracle(S.HILCodec.from_weights(w, 8).cuda(), p, x, 8, o, "config 2")
    print("config 2:", r)
    assert r["frames"] == 4800


def test_config3_full_size_vs_oracle():
    """BASELINE configs[2] (the bench workload): hil_music, 256 x 24000, n_q = 12 -- all 230 400 decisions."""
    cfg, w, _ = _weights("hil_music")
    p = params(w)
    x = synth_wav(256, 24000, seed=1234)      # the bench's rank-0 input
    o = _oracle(cfg, p, x, 12)
    r = _check_against_oracle(S.HILCodec.from_weights(w, 12).cuda(), p, x, 12, o, "config 3")
    print("config 3:", r)
    assert r["frames"] == 19200


def _edge_inputs(T):
    g = torch.Generator().manual_seed(7)
    t = torch.arange(T, dtype=torch.float32)
    impulse0 = torch.zeros(T); impulse0[0] = 1.0
    impulse = torch.zeros(T); impulse[1000] = -1.0
    square = torch.where((t // 37) % 2 == 0, torch.tensor(1.0), torch.tensor(-1.0))
    cases = {
        "zeros": torch.zeros(T),                                   # log(clamp(|STFT|, 1e-5)) floor everywhere
        "plus_full_scale": torch.ones(T),                          # DC at +1.0
        "minus_full_scale": -torch.ones(T),
        "dc_half": torch.full((T,), 0.5),
        "clipped_noise": (3.0 * torch.randn(T, generator=g)).clamp(-1, 1),   # mostly at the rails
        "square_full_scale": square,
        "impulse_t0": impulse0,
        "impulse_t1000": impulse,
        "near_silence": 1e-7 * torch.randn(T, generator=g),       # below the 1e-5 magnitude floor
        "full_scale_sine": torch.sin(2 * np.pi * 440.0 * t / 24000.0),
        "silence_then_burst": torch.cat([torch.zeros(T // 2), 0.5 * torch.randn(T - T // 2, generator=g).clamp(-1, 1)]),
    }
    names = list(cases)
    return names, torch.stack([cases[k] for k in names]).unsqueeze(1)


@pytest.mark.parametrize("name", ["hil_speech", "hil_music"])
def test_edge_case_inputs_vs_oracle(name):
    cfg, w, _ = _weights(name)
    n_q = cfg.num_quantizers
    p = params(w)
    names, x = _edge_inputs(320 * 30)
    o = _oracle(cfg, p, x, n_q)
    m = S.HILCodec.from_weights(w, n_q).cuda()
    r = _check_against_oracle(m, p, x, n_q, o, f"{name} edge cases {names}")
    print(name, "edge cases:", r)
    # frame by frame (hop 320, GPU-resident caches) through the streaming kernels: same decisions under the same policy
    st = m.new_stream_state(x.shape[0])
    ids, ws = [], []
    xd = x.cuda()
    for f in range(30):
        i1, y1 = st.step(xd[:, :, f * 320:(f + 1) * 320], n_q)
        ids.append(i1.clone()); ws.append(y1.clone())
    ids = torch.cat(ids, 2)
    ce, _ = m.initialize_cache(xd)
    z, _ = m.encoder(xd, *ce)
    ill_names = [names[b] for b in r["ill_conditioned_clips"] if isinstance(b, int)]
    good = torch.tensor([b not in r["ill_conditioned_clips"] for b in range(x.shape[0])])
    gi = torch.nonzero(good).flatten()
    bad, worst = index_report(oracle_cfg(n_q), p, z[gi.cuda()], ids[:, gi.cuda()], o["indices"][:, gi], n_q)
    assert bad == 0 or worst < GAP, (bad, worst)
    clean = ~(ids.cpu() != o["indices"]).any(dim=0).any(dim=1)
    assert (torch.cat(ws, 2).cpu() - o["wav"])[clean].abs().max().item() < TOL
    # the ill-conditioned inputs are the DC-like ones, as the CPU analysis predicts (nothing else may hide behind it)
    assert set(ill_names) <= {"plus_full_scale", "minus_full_scale", "dc_half", "full_scale_sine"}, ill_names


def test_interleaved_chunk_sizes_on_one_stream_state():
    """a StreamState serves several (T, n) keys; a longer chunk grows the workspace, which must drop
    the captured graphs of the shorter key instead of replaying them on freed memory."""
    cfg, w, _ = _weights("hil_speech")
    m = S.HILCodec.from_weights(w, 8).cuda()
    x = synth_wav(2, 320 * 64, seed=3).cuda()
    ref_idx, ref_y = m.codec_forward(x, 8)
    st = m.new_stream_state(2)
    sizes = [1, 1, 1, 4, 1, 1, 16, 1, 1, 1, 4, 4, 1, 25, 1, 1]      # hops per call; 1-hop graphs captured, then growth
    pos, ids, ws = 0, [], []
    for h in sizes:
        i1, y1 = st.step(x[:, :, pos:pos + 320 * h], 8)
        ids.append(i1.clone()); ws.append(y1.clone())
        pos += 320 * h
    assert pos == x.shape[2]
    ids, ws = torch.cat(ids, 2), torch.cat(ws, 2)
    assert torch.isfinite(ws).all()
    p = params(w)
    ce, _ = m.initialize_cache(x)
    z, _ = m.encoder(x, *ce)
    bad, worst = index_report(oracle_cfg(8), p, z, ref_idx, ids.cpu(), 8)
    assert bad == 0 or worst < GAP, (bad, worst)
    if bad == 0:
        assert (ws - ref_y).abs().max().item() < 5e-5


def test_stream_state_outlives_weight_reload():
    """replacing the weights destroys the native model; StreamStates created before that must fail
    loudly on use and stay safe to drop, not dereference a freed model."""
    cfg = W.HIL_SPEECH
    m = S.HILCodec.from_weights(W.random_weights(cfg, 1), 8).cuda()
    x = synth_wav(1, 320 * 2, seed=1).cuda()
    st = m.new_stream_state(1)
    st.step(x[:, :, :320], 8)
    m._core.set_weights(W.random_weights(cfg, 2))          # invalidates every native handle
    for call in (lambda: st.step(x[:, :, 320:], 8), st.reset, st.export, lambda: m.codec_forward(x, 8, state=st)):
        with pytest.raises(RuntimeError):
            call()
    del st                                                  # hil_state_destroy on a state whose model is gone
    st2 = m.new_stream_state(1)
    i2, _ = st2.step(x[:, :, :320], 8)
    i3, _ = m.codec_forward(x[:, :, :320], 8)
    assert torch.equal(i2, i3)


def test_wrong_dtype_and_shape_are_rejected():
    m = S.HILCodec.from_weights(W.random_weights(W.HIL_SPEECH, 1), 8).cuda()
    x = synth_wav(1, 640).cuda()
    with pytest.raises(TypeError):
        m.codec_forward(x.long(), 8)
    with pytest.raises(TypeError):
        m.codec_forward(x.double(), 8)
    with pytest.raises(ValueError):
        m.codec_forward(x[0], 8)
    with pytest.raises(ValueError):
        m.codec_forward(x.repeat(1, 2, 1), 8)
    with pytest.raises(TypeError):
        m.dequantizer(torch.zeros(8, 1, 2, device="cuda"), 8)      # float indices
    with pytest.raises(TypeError):
        m.encoder(x, *[c.long() for c in m.encoder.initialize_cache(x)])


def test_fp16_range_guard_reroutes_to_fp32_kernels():
    """the tensor-core kernels split activations into fp16 pairs and need |x| < 65504; the reference has no
    such limit.  Audio scaled by 1e6 overflows the split in the first ResBlock: unguarded, the latents come out NaN and
    the state's range flag is set; guarded (the default), every entry point repeats the call on the FP32 kernels and
    matches the oracle under the usual bars."""
    import ctypes as C

    from hilcodec_b200 import _lib

    cfg, w, _ = _weights("hil_speech")
    p = params(w)
    x = synth_wav(2, 320 * 20, seed=5) * 1.0e6
    o = _oracle(cfg, p, x, 8)
    assert torch.isfinite(o["wav"]).all() and torch.isfinite(o["z"]).all()     # fine in the reference's fp32
    m = S.HILCodec.from_weights(w, 8).cuda()
    core, dev = m._core, torch.device("cuda", torch.cuda.current_device())
    lib = _lib.load()
    xd = x.cuda()

    core.check_range = False
    ce, _ = m.initialize_cache(xd)
    z0, _ = m.encoder(xd, *ce)
    assert not torch.isfinite(z0).all()                                        # the hole, unguarded
    flag = C.c_int32(0)
    _lib.check(lib.hil_state_range_flag(core.state(dev, 2), 1, torch.cuda.current_stream().cuda_stream, C.byref(flag)))
    assert flag.value == 1
    st = m.new_stream_state(2)
    st.step(xd[:, :, :320], 8)
    assert st.range_overflow() and not st.range_overflow()                     # sticky until cleared

    core.check_range = True
    r = _check_against_oracle(m, p, x, 8, o, "range guard")
    print("range guard:", r)

    # the C-ABI host call guards itself (include/hilcodec_b200.h)
    F = x.shape[2] // 320
    xh = x.contiguous().pin_memory()
    ih = torch.empty(8, 2, F, dtype=torch.int64).pin_memory()
    yh = torch.empty(2, 1, x.shape[2]).pin_memory()
    state = core.state(dev, 2)
    _lib.check(lib.hil_state_reset(state, torch.cuda.current_stream().cuda_stream))
    _lib.check(lib.hil_codec_forward_host(core.model(dev), state, xh.data_ptr(), 2, x.shape[2], 8, ih.data_ptr(),
                                          yh.data_ptr(), torch.cuda.current_stream().cuda_stream))
    idx, y = m.codec_forward(xd, 8)
    assert torch.equal(ih, idx.cpu()) and torch.equal(yh, y.cpu())
    # in-range audio never takes the detour: same bits with the guard on and off
    x1 = synth_wav(2, 320 * 20, seed=6).cuda()
    a = m.codec_forward(x1, 8)
    core.check_range = False
    b = m.codec_forward(x1, 8)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
