"""SURVEY.md section 8f rows on the GPU: bitstream packing, the test_onnx.py-compatible runner, the
training-graph call signatures."""
import os

import numpy as np
import pytest
import torch

from hilcodec_b200 import bitstream, models, onnx_runner
from hilcodec_b200 import streaming as S
from hilcodec_b200 import weights as W
from oracle import bitstream_oracle
from oracle import hilcodec_oracle as O

from helpers import GOLDEN, params, synth_wav

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_q,n", [(8, 8), (12, 12), (12, 3), (8, 1)])
def test_bitstream_roundtrip_and_format(n_q, n):
    cfg = W.CodecConfig(num_quantizers=n_q)
    m = S.HILCodec.from_weights(W.random_weights(cfg, 1), n_q).cuda()
    g = torch.Generator().manual_seed(n)
    idx = torch.randint(0, 1024, (n, 3, 77), generator=g)
    idx[:, 0, 0] = 1023
    idx[:, 0, 1] = 0
    packed = bitstream.pack(m, idx.cuda())
    assert packed.shape == (3, 77, (10 * n + 7) // 8) and packed.dtype == torch.uint8
    assert np.array_equal(packed.cpu().numpy(), bitstream_oracle.pack_numpy(idx.numpy()))
    back = bitstream.unpack(m, packed, n)
    assert torch.equal(back.cpu(), idx)
    assert np.array_equal(bitstream_oracle.unpack_numpy(packed.cpu().numpy(), n), idx.numpy())
    # known answer: two 10-bit values 0x3FF, 0x001 -> bytes FF 07 00 (LSB first)
    if n >= 2:
        kat = np.array([[[0x3FF]], [[0x001]]] + [[[0]]] * (n - 2))
        assert bitstream_oracle.pack_numpy(kat)[0, 0, :3].tolist() == [0xFF, 0x07, 0x00]


@pytest.mark.skipif(not W.have_pretrained("hil_speech"), reason="published weights not extracted")
def test_onnx_runner_reproduces_reference_artefacts(tmp_path):
    """The reference runner's workflow on 2 s of its own clip: frame-by-frame encode -> int16 [n,B,T] .npy ->
    decode 5 frames per call -> wav, compared with the reference's committed outputs."""
    from scipy.io import wavfile

    g = np.load(os.path.join(GOLDEN, "speech_kat.npz"))
    frames = 150
    wav_path = os.path.join(str(tmp_path), "in.wav")
    wavfile.write(wav_path, 24000, g["wav_in"][:frames * 320 + 100])  # ragged tail is cut to a hop multiple
    rc = onnx_runner.main(["-n", "hil_speech", "-q", "8", "-f", "5", "--enc", "--dec", "--input", wav_path,
                           "--outdir", str(tmp_path)])
    assert rc == 0
    q = np.load(os.path.join(str(tmp_path), "hil_speech_quantized.npy"))
    assert q.dtype == np.int16 and q.shape == (8, 1, frames)
    assert np.array_equal(q, g["indices"][:, :, :frames])
    sr, out = wavfile.read(os.path.join(str(tmp_path), "hil_speech_output.wav"))
    assert sr == 24000 and out.shape == (frames * 320,)
    assert np.abs(out.astype(np.int32) - g["wav_out"][:frames * 320].astype(np.int32)).max() <= 2


@pytest.mark.skipif(not W.have_pretrained("hil_speech"), reason="published weights not extracted")
def test_onnx_runner_cache_npz_protocol(tmp_path, capsys):
    """`e_in{i}` / `d_in{i}` dicts seeded from `{name}_cache_{enc,dec}.npz` (test_onnx.py:73, 80-81, 121, 133-134): the
    zero caches the runner writes have the reference's names and shapes, a run seeded from them equals a run without
    the files, and NON-zero caches (the state after the first half of a clip) resume the stream exactly."""
    from scipy.io import wavfile

    g = np.load(os.path.join(GOLDEN, "speech_kat.npz"))
    d = str(tmp_path)
    assert onnx_runner.main(["-n", "hil_speech", "--save-cache", "--outdir", d]) == 0
    ce, cd = np.load(os.path.join(d, "hil_speech_cache_enc.npz")), np.load(os.path.join(d, "hil_speech_cache_dec.npz"))
    assert ce.files == [f"e_in{i}" for i in range(22)] and cd.files == [f"d_in{i}" for i in range(30)]
    assert ce["e_in0"].shape == (1, 1, 1023) and cd["d_in0"].shape == (1, 1536, 4) and not ce["e_in5"].any()
    frames = 40
    wavfile.write(os.path.join(d, "a.wav"), 24000, g["wav_in"][:frames * 320])
    wavfile.write(os.path.join(d, "b.wav"), 24000, g["wav_in"][frames * 320:2 * frames * 320])
    assert onnx_runner.main(["-n", "hil_speech", "-q", "8", "-t", "2", "--enc", "--dec", "--input", os.path.join(d, "a.wav"),
                             "--outdir", d]) == 0
    out = capsys.readouterr().out
    assert "wav length: " in out and "encoder: " in out and "decoder: " in out and "rtf: " in out and "(↑)" in out
    qa = np.load(os.path.join(d, "hil_speech_quantized.npy"))
    assert np.array_equal(qa, g["indices"][:, :, :frames])
    # second half, seeded with the caches left by the first half
    m = S.HILCodec.from_pretrained("hil_speech").cuda()
    x = torch.from_numpy(g["wav_in"][:frames * 320].astype(np.float32) / 32768).view(1, 1, -1).cuda()
    e, dcache = m.initialize_cache(x)
    z, e = m.encoder(x, *e)
    _, dcache = m.decoder(m.dequantizer(m.quantizer(z, 8), 8), *dcache)
    d2 = os.path.join(d, "resume")
    os.makedirs(d2)
    np.savez(os.path.join(d2, "hil_speech_cache_enc.npz"), **{f"e_in{i}": c.cpu().numpy() for i, c in enumerate(e)})
    np.savez(os.path.join(d2, "hil_speech_cache_dec.npz"), **{f"d_in{i}": c.cpu().numpy() for i, c in enumerate(dcache)})
    assert onnx_runner.main(["-n", "hil_speech", "-q", "8", "--enc", "--dec", "--input", os.path.join(d, "b.wav"),
                             "--outdir", d2, "--cache-dir", d2]) == 0
    qb = np.load(os.path.join(d2, "hil_speech_quantized.npy"))
    assert np.array_equal(qb, g["indices"][:, :, frames:2 * frames])
    _, outb = wavfile.read(os.path.join(d2, "hil_speech_output.wav"))
    assert np.abs(outb.astype(np.int32) - g["wav_out"][frames * 320:2 * frames * 320].astype(np.int32)).max() <= 2
    # wrong shapes / missing names are errors, not silent zeros
    np.savez(os.path.join(d2, "hil_speech_cache_enc.npz"), e_in0=np.zeros((1, 1, 5), np.float32))
    with pytest.raises(KeyError):
        onnx_runner.main(["-n", "hil_speech", "--enc", "--input", os.path.join(d, "b.wav"), "--outdir", d2])


def test_training_graph_signatures():
    cfg = W.HIL_MUSIC
    w = W.random_weights(cfg, 6)
    m = models.HILCodec(S.HILCodec.from_weights(w, 12).cuda())
    x = synth_wav(2, 320 * 20, seed=3)
    wav, num_replaces, loss = m(x.cuda(), 5)
    assert wav.shape == (2, 1, 6400) and wav.dtype == torch.float32
    assert num_replaces.shape == (12,) and num_replaces.dtype == np.int64 and not num_replaces.any()
    z = m.encode(x.cuda())
    assert z.shape == (2, 128, 20)
    q, _, loss2, idx = m.quantizer(z, 5, return_indices=True)
    assert q.shape == (2, 128, 20) and idx.shape == (2, 5, 20)
    p = params(w)
    with torch.no_grad():
        o = O.codec_forward(O.CodecConfig(num_quantizers=12), p, x, 5)
    assert torch.equal(idx.permute(1, 0, 2).cpu(), o["indices"])
    assert (wav.cpu() - o["wav"]).abs().max().item() < 1e-4
    ref_loss = torch.nn.functional.mse_loss(o["z"].transpose(1, 2), o["q"].transpose(1, 2))
    assert abs(float(loss) - float(ref_loss)) < 1e-5 and abs(float(loss2) - float(ref_loss)) < 1e-5
    with pytest.raises(AssertionError):
        m.quantizer(z, 13)
