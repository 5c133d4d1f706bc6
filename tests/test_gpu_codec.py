"""Whole-path parity (-m gpu): the CUDA encode -> RVQ -> decode through the drop-in modules
(which call the C ABI) against the CPU oracle, the committed reference fixtures and the
reference's golden clip.

Bars (BASELINE.json north_star): VQ indices bit-exact; latents and decoded PCM max-abs-err
< 1e-4.  Near-tie policy for large batches: a disagreement is tolerated only if the fp64
relative gap between the two best codes at that decision is < 1e-5."""
import os

import numpy as np
import pytest
import torch

from hilcodec_b200 import streaming as S
from hilcodec_b200 import weights as W
from oracle import hilcodec_oracle as O

from helpers import GOLDEN, index_report, oracle_cfg, params, synth_wav

pytestmark = pytest.mark.gpu

TOL = 1e-4


def _model(w, n_q):
    return S.HILCodec.from_weights(w, n_q).cuda()


@pytest.mark.parametrize("name", ["ref_random_speech.npz", "ref_random_music.npz"])
def test_reference_fixture_four_call_flow(name):
    """Committed outputs of the reference's own streaming.py classes (random weights)."""
    g = np.load(os.path.join(GOLDEN, name))
    n_q = int(g["n_q"])
    w = W.random_weights(W.CodecConfig(num_quantizers=n_q), int(g["seed"]))
    m = _model(w, n_q)
    x = torch.from_numpy(g["x"]).cuda()
    ce, cd = m.initialize_cache(x)
    assert len(ce) == 22 and len(cd) == 30
    z, ce = m.encoder(x, *ce)
    idx = m.quantizer(z, n_q)
    q = m.dequantizer(idx, n_q)
    y, cd = m.decoder(q, *cd)
    torch.cuda.synchronize()
    assert z.shape == g["z"].shape and y.shape == g["wav"].shape and idx.dtype == torch.int64
    assert np.abs(z.cpu().numpy() - g["z"]).max() < TOL
    assert np.array_equal(idx.cpu().numpy(), g["indices"].astype(np.int64))
    assert np.array_equal(q.cpu().numpy(), g["q"])  # gathers + in-order adds: bit-exact
    assert np.abs(y.cpu().numpy() - g["wav"]).max() < TOL
    for i, c in enumerate(ce):
        assert tuple(c.shape) == g[f"enc_cache{i}"].shape
        assert np.abs(c.cpu().numpy() - g[f"enc_cache{i}"]).max() < TOL, f"enc cache {i}"
    for i, c in enumerate(cd):
        assert tuple(c.shape) == g[f"dec_cache{i}"].shape
        assert np.abs(c.cpu().numpy() - g[f"dec_cache{i}"]).max() < TOL, f"dec cache {i}"


@pytest.mark.parametrize("name", ["ref_random_speech.npz", "ref_random_music.npz"])
def test_reference_fixture_streaming(name):
    """Chunked calls with the caches handed back and forth as a list of tensors, exactly like
    scripts/HILCodec Onnx.ipynb cell 3 / test_onnx.py:75-93."""
    g = np.load(os.path.join(GOLDEN, name))
    n_q, hops = int(g["n_q"]), int(g["stream_hops"])
    w = W.random_weights(W.CodecConfig(num_quantizers=n_q), int(g["seed"]))
    m = _model(w, n_q)
    x = torch.from_numpy(g["x"]).cuda()
    ce, cd = m.initialize_cache(x)
    ids, ws = [], []
    step = hops * 320
    for s in range(0, x.shape[2], step):
        z, ce = m.encoder(x[:, :, s:s + step], *ce)
        idx = m.quantizer(z, n_q)
        y, cd = m.decoder(m.dequantizer(idx, n_q), *cd)
        ids.append(idx)
        ws.append(y)
    idx = torch.cat(ids, 2).cpu().numpy()
    y = torch.cat(ws, 2).cpu().numpy()
    assert np.array_equal(idx, g["stream_indices"].astype(np.int64))
    assert np.abs(y - g["stream_wav"]).max() < TOL


def test_fused_forward_and_stateful_streaming_match_four_call_flow():
    n_q = 12
    w = W.random_weights(W.HIL_MUSIC, 5)
    m = _model(w, n_q)
    x = synth_wav(3, 320 * 9, seed=11).cuda()
    ce, cd = m.initialize_cache(x)
    z, _ = m.encoder(x, *ce)
    idx = m.quantizer(z, n_q)
    y, _ = m.decoder(m.dequantizer(idx, n_q), *cd)
    idx2, y2 = m.codec_forward(x, n_q)
    assert torch.equal(idx, idx2) and torch.equal(y, y2)
    # GPU-resident state, frame by frame (hop 320), must reproduce the one-shot result
    st = m.new_stream_state(3)
    ids, ws = [], []
    for s in range(0, x.shape[2], 320):
        i3, y3 = m.codec_forward(x[:, :, s:s + 320], n_q, state=st)
        ids.append(i3)
        ws.append(y3)
    assert torch.equal(torch.cat(ids, 2), idx)
    # one-frame chunks run the short-T layers on the FFMA kernels and one-shot runs them on the
    # tensor-core kernels: same math, different fp32 summation order
    assert (torch.cat(ws, 2) - y).abs().max().item() < 5e-5
    # state export -> list of tensors -> import round trip is lossless
    e1, d1 = st.export()
    st2 = m.new_stream_state(3)
    st2.load(e1, d1)
    e2, d2 = st2.export()
    assert all(torch.equal(a, b) for a, b in zip(e1 + d1, e2 + d2))
    # and equals the caches the functional protocol returns after the same audio
    ce, cd = m.initialize_cache(x)
    _, ce = m.encoder(x, *ce)
    for a, b in zip(e1, ce):
        assert (a - b).abs().max().item() < 5e-5


@pytest.mark.parametrize("cfg_name,B,frames", [("hil_speech", 4, 75), ("hil_music", 2, 75)])
def test_oracle_parity_random_weights(cfg_name, B, frames):
    cfg = W.CONFIGS[cfg_name]
    n_q = cfg.num_quantizers
    w = W.random_weights(cfg, 21)
    m = _model(w, n_q)
    x = synth_wav(B, 320 * frames, seed=77)
    p = params(w)
    with torch.no_grad():
        o = O.codec_forward(oracle_cfg(n_q), p, x, n_q)
    xd = x.cuda()
    ce, cd = m.initialize_cache(xd)
    z, _ = m.encoder(xd, *ce)
    idx, q = m.quantizer.quantize(z, n_q)
    y, _ = m.decoder(q, *cd)
    assert (z.cpu() - o["z"]).abs().max().item() < TOL
    bad, worst = index_report(oracle_cfg(n_q), p, z, idx, o["indices"], n_q)
    assert bad == 0 or worst < 1e-5, (bad, worst)
    if bad == 0:
        assert (y.cpu() - o["wav"]).abs().max().item() < TOL
    # decoder alone on the oracle's q: isolates decoder parity from index near-ties
    y2, _ = m.decoder(o["q"].cuda(), *cd)
    assert (y2.cpu() - o["wav"]).abs().max().item() < TOL
    # n < num_quantizers and the assert on n (vector_quantize.py:213)
    idx3 = m.quantizer(z, 3)
    assert torch.equal(idx3, idx[:3])
    with pytest.raises(AssertionError):
        m.quantizer(z, n_q + 1)


def test_rvq_bit_exact_given_identical_latents():
    """Codebook search + residual + dequant on latents shared with the oracle."""
    cfg = W.HIL_MUSIC
    w = W.random_weights(cfg, 2)
    p = params(w)
    m = _model(w, 12)
    g = torch.Generator().manual_seed(3)
    z = torch.randn(5, 301, 128, generator=g)
    z = z / z.norm(dim=2, keepdim=True) * 128 ** 0.5
    ref = O.rvq_encode(oracle_cfg(12), p, z, 12)
    idx, qsum = m.quantizer.quantize(z.cuda(), 12)
    bad, worst = index_report(oracle_cfg(12), p, z, idx, ref, 12)
    assert bad == 0 or worst < 1e-6, (bad, worst)
    q_ref = O.rvq_decode(oracle_cfg(12), p, idx.cpu(), 12)
    assert torch.equal(m.dequantizer(idx, 12).cpu(), q_ref)
    assert torch.equal(qsum.cpu(), q_ref)


@pytest.mark.skipif(not W.have_pretrained("hil_speech"), reason="published weights not extracted")
def test_golden_clip_full():
    """The reference's KAT, all 30.6 s: 18 368 / 18 368 indices, PCM within one int16 LSB."""
    g = np.load(os.path.join(GOLDEN, "speech_kat.npz"))
    m = S.HILCodec.from_pretrained("hil_speech").cuda()
    x = torch.from_numpy(g["wav_in"].astype(np.float32) / 32768).view(1, 1, -1).cuda()
    idx, y = m.codec_forward(x, 8)
    gold = torch.from_numpy(g["indices"].astype(np.int64))
    assert idx.shape == gold.shape
    assert int((idx.cpu() != gold).sum()) == 0
    ref = g["wav_out"].astype(np.float32) / 32768
    # the golden file is int16 PCM: one LSB of quantisation (3.05e-5) plus our own error budget
    assert np.abs(y[0, 0].cpu().numpy() - ref).max() <= 1.0 / 32768 + 2e-5


@pytest.mark.skipif(not W.have_pretrained("hil_music"), reason="published weights not extracted")
def test_pretrained_music_vs_oracle():
    cfg = W.HIL_MUSIC
    w = W.load_pretrained("hil_music")
    m = S.HILCodec.from_pretrained("hil_music").cuda()
    x = synth_wav(2, 24000, seed=1234)
    p = params(w)
    with torch.no_grad():
        o = O.codec_forward(oracle_cfg(12), p, x, 12)
    xd = x.cuda()
    ce, cd = m.initialize_cache(xd)
    z, _ = m.encoder(xd, *ce)
    idx, q = m.quantizer.quantize(z, 12)
    y, _ = m.decoder(q, *cd)
    assert (z.cpu() - o["z"]).abs().max().item() < TOL
    bad, worst = index_report(oracle_cfg(12), p, z, idx, o["indices"], 12)
    assert bad == 0 or worst < 1e-5, (bad, worst)
    y2, _ = m.decoder(o["q"].cuda(), *cd)
    assert (y2.cpu() - o["wav"]).abs().max().item() < TOL


@pytest.mark.skipif(not W.have_pretrained("hil_music"), reason="published hil_music weights did not travel to this box")
def test_pretrained_music_reference_fixture():
    """The reference's own classes with the published hil_music weights on real speech (tests/golden/make_golden.py
    music): the only reference-made pin `hil_music` has."""
    g = np.load(os.path.join(GOLDEN, "ref_music_published.npz"))
    w = W.load_pretrained("hil_music")
    m = _model(w, 12)
    x = torch.from_numpy(g["x"]).cuda()
    ce, cd = m.initialize_cache(x)
    z, _ = m.encoder(x, *ce)
    idx = m.quantizer(z, 12)
    y, _ = m.decoder(m.dequantizer(idx, 12), *cd)
    assert np.abs(z.cpu().numpy() - g["z"]).max() < TOL
    ref_idx = torch.from_numpy(g["indices"].astype(np.int64))
    bad, worst = index_report(oracle_cfg(12), params(w), z, idx, ref_idx, 12)
    assert bad == 0 or worst < 1e-5, (bad, worst)
    if bad == 0:
        assert np.abs(y.cpu().numpy() - g["wav"]).max() < TOL


def test_full_size_properties_config2():
    """BASELINE config 2 size (64 x 24000): too big for the CPU oracle in a test, so check
    size-independent properties: batch rows are independent (row b of the batch == the same
    clip run alone), chunked == one-shot, outputs finite, ||z_t|| = sqrt(128)."""
    cfg = W.HIL_SPEECH
    w = W.load_pretrained("hil_speech") if W.have_pretrained("hil_speech") else W.random_weights(cfg, 4)
    m = _model(w, 8)
    x = synth_wav(64, 24000, seed=1234).cuda()
    idx, y = m.codec_forward(x, 8)
    assert torch.isfinite(y).all() and y.abs().max().item() <= 1.0
    for b in (0, 17, 63):
        i1, y1 = m.codec_forward(x[b:b + 1], 8)
        assert torch.equal(i1[:, 0], idx[:, b])
        assert (y1[0] - y[b]).abs().max().item() < 1e-5
    ce, _ = m.initialize_cache(x)
    z, _ = m.encoder(x, *ce)
    assert (z.norm(dim=2) - 128 ** 0.5).abs().max().item() < 1e-3
    st = m.new_stream_state(64)
    ids = []
    for s in range(0, 24000, 8000):  # 25-frame chunks
        i2, _ = m.codec_forward(x[:, :, s:s + 8000], 8, state=st)
        ids.append(i2)
    # chunked and one-shot runs route some layers through different kernels (FFMA for short
    # chunks, tensor cores otherwise), so a handful of near-tie decisions may differ: every
    # differing frame must be a near tie (fp64 gap of the two best codes < 1e-5)
    idc = torch.cat(ids, 2)
    same = (idc == idx).float().mean().item()
    assert same > 0.999, same
    p = params(w)
    bad, worst = index_report(oracle_cfg(8), p, z, idx, idc.cpu(), 8)
    assert bad == 0 or worst < 1e-5, (bad, worst)


@pytest.mark.parametrize("name,n_q,streams", [("hil_music", 12, 64), ("hil_speech", 8, 40)])
def test_many_streams_hop_by_hop_matches_one_shot(name, n_q, streams):
    """Config 4 with many concurrent streams, one 320-sample hop per call.  At >= 16 streams the 8-sample layers of a hop
    (the widest of the model) run on flat tensor-core tiles (128 / 8 whole clips per tile: gemm_h.cu `flat_t`), the
    40-window STFT on the tensor-core STFT kernel, the short depthwise rows many per CTA.  The hop-by-hop result must be
    the one-shot result of the same audio: indices equal up to fp64 near ties, PCM within the bar on every stream whose
    indices all agree."""
    cfg = W.CONFIGS[name]
    w = W.load_pretrained(name) if W.have_pretrained(name) else W.random_weights(cfg, 4)
    m = _model(w, n_q)
    hops = 8
    x = synth_wav(streams, 320 * hops, seed=77).cuda()
    idx, y = m.codec_forward(x, n_q)
    ce, _ = m.initialize_cache(x)
    z, _ = m.encoder(x, *ce)
    st = m.new_stream_state(streams)
    ids, ys = [], []
    for h in range(hops):
        i2, y2 = m.codec_forward(x[:, :, 320 * h:320 * (h + 1)], n_q, state=st)
        ids.append(i2)
        ys.append(y2)
    idc, yc = torch.cat(ids, 2), torch.cat(ys, 2)
    assert torch.isfinite(yc).all()
    assert (idc == idx).float().mean().item() > 0.995
    bad, worst = index_report(oracle_cfg(n_q), params(w), z, idx, idc.cpu(), n_q)
    assert bad == 0 or worst < 1e-5, (bad, worst)
    rows_equal = (idc == idx).all(dim=0).all(dim=1)          # streams whose every index agrees
    assert rows_equal.float().mean().item() > 0.8
    assert (yc[rows_equal] - y[rows_equal]).abs().max().item() < TOL


def test_full_size_properties_config3():
    """BASELINE config 3 size (hil_music, 256 x 24000, n_q = 12: the bench workload, 19 200 frames per step): batch
    rows are independent of the batch they ride in, outputs are finite and bounded by tanh, ||z_t|| = sqrt(128),
    and dequantising the indices reproduces the decoder input."""
    cfg = W.HIL_MUSIC
    w = W.load_pretrained("hil_music") if W.have_pretrained("hil_music") else W.random_weights(cfg, 4)
    m = _model(w, 12)
    x = synth_wav(256, 24000, seed=4321).cuda()
    idx, y = m.codec_forward(x, 12)
    assert idx.shape == (12, 256, 75) and y.shape == (256, 1, 24000)
    assert torch.isfinite(y).all() and y.abs().max().item() <= 1.0
    assert int(idx.min()) >= 0 and int(idx.max()) < 1024
    p = params(w)
    for b in (0, 101, 255):
        xb = x[b:b + 1].contiguous()
        i1, y1 = m.codec_forward(xb, 12)
        ce, _ = m.initialize_cache(xb)
        zb, _ = m.encoder(xb, *ce)
        bad, worst = index_report(oracle_cfg(12), p, zb, i1, idx[:, b:b + 1].cpu(), 12)
        assert bad == 0 or worst < 1e-5, (b, bad, worst)
        if bad == 0:
            assert (y1[0] - y[b]).abs().max().item() < 1e-4
    ce, cd = m.initialize_cache(x)
    z, _ = m.encoder(x, *ce)
    assert (z.norm(dim=2) - 128 ** 0.5).abs().max().item() < 1e-3
    y2, _ = m.decoder(m.dequantizer(idx, 12), *cd)   # the four-call flow's second half on the fused path's indices
    assert (y2 - y).abs().max().item() < 1e-4
    del x, y, y2, z


def test_config4_streaming_30s_clip():
    """BASELINE config 4: hil_music, one 30 s stream fed hop by hop (hop = 320: 2250 sequential frames; the
    config's "hop=300" is AudioDec's hop and cannot be fed to HILCodec) with the per-layer
    caches resident on the GPU, against the one-shot result of the same clip."""
    cfg = W.HIL_MUSIC
    w = W.load_pretrained("hil_music") if W.have_pretrained("hil_music") else W.random_weights(cfg, 8)
    m = _model(w, 12)
    frames = 2250
    x = synth_wav(1, 320 * frames, seed=30).cuda()
    idx, y = m.codec_forward(x, 12)
    st = m.new_stream_state(1)
    ids, ws = [], []
    for f in range(frames):
        i1, y1 = m.codec_forward(x[:, :, f * 320:(f + 1) * 320], 12, state=st)
        ids.append(i1)
        ws.append(y1)
    ids = torch.cat(ids, 2)
    ws = torch.cat(ws, 2)
    same = (ids == idx).float().mean().item()
    assert same > 0.999, same
    p = params(w)
    ce, _ = m.initialize_cache(x)
    z, _ = m.encoder(x, *ce)
    bad, worst = index_report(oracle_cfg(12), p, z, idx, ids.cpu(), 12)
    assert bad == 0 or worst < 1e-5, (bad, worst)
    if bad == 0:
        assert (ws - y).abs().max().item() < 5e-5


def test_graph_streaming_executor_matches_eager():
    """CUDA-graph replay of the per-frame step (hil_codec_forward_graph) == the eager stateful path."""
    n_q = 12
    w = W.random_weights(W.HIL_MUSIC, 12)
    m = _model(w, n_q)
    x = synth_wav(2, 320 * 40, seed=5).cuda()
    st_e = m.new_stream_state(2)
    st_g = m.new_stream_state(2)
    for f in range(40):
        chunk = x[:, :, f * 320:(f + 1) * 320]
        ie, ye = m.codec_forward(chunk, n_q, state=st_e)
        ig, yg = st_g.step(chunk, n_q)
        assert torch.equal(ie, ig), f
        assert torch.equal(ye, yg), f
    ee, de = st_e.export()
    eg, dg = st_g.export()
    assert all(torch.equal(a, b) for a, b in zip(ee + de, eg + dg))
