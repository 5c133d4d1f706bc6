"""Generate the committed golden fixtures (run in the dev container, where
/root/reference exists):

    python tests/golden/make_golden.py

* `speech_kat.npz`   -- the reference's only known-answer test, copied as data:
  `onnx/input_speech.wav` (first 2296*320 samples, int16), `onnx/hil_speech_quantized.npy`
  (int16 [8,1,2296], written by test_onnx.py:100) and `onnx/hil_speech_output.wav`
  (int16, test_onnx.py:139).  Needs the published hil_speech weights at run time.
* `ref_random_*.npz` -- outputs of the reference's OWN `models/hilcodec/streaming.py`
  classes (imported through oracle/ref_shim.py) on seeded random weights
  (`hilcodec_b200.weights.random_weights(cfg, seed)`, reproducible anywhere), one-shot
  and frame-by-frame, so the CUDA path can be checked on the GPU box where the reference
  itself cannot be imported.
* `ref_train_random.npz` -- outputs of the reference's TRAINING graph (`models/hilcodec/models.py`
  `HILCodec.forward`, eval) on a seeded training-format checkpoint
  (`hilcodec_b200.checkpoint.random_training_state_dict(cfg, seed)`: weight-norm pairs, un-merged
  scales), for input lengths that are and are not multiples of the hop.

    python tests/golden/make_golden.py train     # only (re)generate the training-graph fixture
* `ref_music_published.npz` -- the reference's own streaming.py classes with the PUBLISHED hil_music weights on
  60 frames of onnx/input_speech.wav (n_q = 12): hil_music has no golden output in the reference, this is its pin.

    python tests/golden/make_golden.py music
* `ref_generic_rvq.npz` -- the reference's GENERIC quantizer (`modules/vector_quantize.py:471 ResidualVQ`, the class
  north_star names) in eval mode on seeded codebooks and latents, channel-first and channel-last, n = None / 3 / 1.

    python tests/golden/make_golden.py generic_rvq
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from hilcodec_b200 import weights as W  # noqa: E402
from oracle import ref_shim  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def synth_wav(batch, samples, seed):
    g = torch.Generator().manual_seed(seed)
    return (0.1 * torch.randn(batch, 1, samples, generator=g)).clamp(-1, 1)


def speech_kat():
    from scipy.io import wavfile

    onnx = os.path.join(ref_shim.REF, "onnx")
    _, wav = wavfile.read(os.path.join(onnx, "input_speech.wav"))
    gold = np.load(os.path.join(onnx, "hil_speech_quantized.npy"))
    _, out = wavfile.read(os.path.join(onnx, "hil_speech_output.wav"))
    frames = gold.shape[2]
    np.savez_compressed(os.path.join(HERE, "speech_kat.npz"),
                        wav_in=wav[:frames * 320].astype(np.int16),
                        indices=gold.astype(np.int16),
                        wav_out=out.astype(np.int16))


def ref_random(name, n_q, seed, batch, frames, stream_hops):
    cfg = W.CodecConfig(num_quantizers=n_q)
    w = W.random_weights(cfg, seed)
    model = ref_shim.build_reference_model(w, n_q)
    x = synth_wav(batch, frames * cfg.hop, 1234 + seed)
    one = ref_shim.reference_forward(model, x, n_q)
    out = {
        "seed": np.int64(seed), "n_q": np.int64(n_q), "x": x.numpy(),
        "z": one["z"].numpy(), "indices": one["indices"].numpy().astype(np.int16),
        "q": one["q"].numpy(), "wav": one["wav"].numpy(),
    }
    for i, c in enumerate(one["enc_caches"]):
        out[f"enc_cache{i}"] = c.numpy()
    for i, c in enumerate(one["dec_caches"]):
        out[f"dec_cache{i}"] = c.numpy()
    # chunked streaming run (chunks of `stream_hops` hops): outputs must equal one-shot ones
    ce, cd = model.initialize_cache(x)
    zs, ids, ws = [], [], []
    step = stream_hops * cfg.hop
    for s in range(0, x.shape[2], step):
        r = ref_shim.reference_forward(model, x[:, :, s:s + step], n_q, ce, cd)
        ce, cd = r["enc_caches"], r["dec_caches"]
        zs.append(r["z"]); ids.append(r["indices"]); ws.append(r["wav"])
    out["stream_hops"] = np.int64(stream_hops)
    out["stream_z"] = torch.cat(zs, 1).numpy()
    out["stream_indices"] = torch.cat(ids, 2).numpy().astype(np.int16)
    out["stream_wav"] = torch.cat(ws, 2).numpy()
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, "stream==oneshot idx", np.array_equal(out["stream_indices"], out["indices"]),
          "wav", float(np.abs(out["stream_wav"] - out["wav"]).max()))


def ref_music_published(name, frames=60):
    """`hil_music` has no golden output in the reference: pin it with the reference's OWN classes run here on the
    PUBLISHED weights (extracted from onnx/hil_music_*.onnx) and real speech (the head of onnx/input_speech.wav)."""
    from scipy.io import wavfile

    w = W.load_pretrained("hil_music")
    model = ref_shim.build_reference_model(w, 12)
    _, wav = wavfile.read(os.path.join(ref_shim.REF, "onnx", "input_speech.wav"))
    x = torch.from_numpy(wav[24000:24000 + frames * 320].astype(np.float32) / 32768.0).view(1, 1, -1)
    r = ref_shim.reference_forward(model, x, 12)
    np.savez_compressed(os.path.join(HERE, name), x=x.numpy(), z=r["z"].numpy(),
                        indices=r["indices"].numpy().astype(np.int16), wav=r["wav"].numpy())
    print(name, tuple(r["indices"].shape), float(r["wav"].abs().max()))


def ref_train(name, n_q, seed, n, batch, lengths):
    from hilcodec_b200 import checkpoint

    cfg = W.CodecConfig(num_quantizers=n_q)
    model = ref_shim.build_reference_training_model(checkpoint.random_training_state_dict(cfg, seed), n_q)
    out = {"seed": np.int64(seed), "n_q": np.int64(n_q), "n": np.int64(n), "lengths": np.asarray(lengths, np.int64)}
    for T in lengths:
        x = synth_wav(batch, T, 4321 + T)
        r = ref_shim.reference_training_forward(model, x, n)
        out[f"x_{T}"] = x.numpy()
        out[f"z_{T}"] = r["z"].numpy()
        out[f"q_{T}"] = r["q"].numpy()
        out[f"indices_{T}"] = r["indices"].numpy().astype(np.int16)
        out[f"wav_{T}"] = r["wav"].numpy()
        out[f"loss_{T}"] = np.float32(r["loss_vq"].item())
        print(name, T, tuple(r["z"].shape), tuple(r["wav"].shape), float(r["loss_vq"]))
    np.savez_compressed(os.path.join(HERE, name), **out)


def generic_rvq_inputs(seed, n_q, size, batch, frames):
    """Seeded codebooks and latents of the generic-RVQ fixture (numpy PCG64: reproducible on the GPU box)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    embeds = [(rng.standard_normal((size, 128)) * (0.8 ** i)).astype(np.float32) for i in range(n_q)]
    x = rng.standard_normal((batch, 128, frames)).astype(np.float32)
    x *= np.float32(11.3137) / np.sqrt((x ** 2).sum(1, keepdims=True))   # like the encoder's L2-normalised latents
    return embeds, x


def ref_generic_rvq(name, seed=11, n_q=5, size=1024, batch=3, frames=50):
    ref_shim.import_streaming()
    from modules.vector_quantize import ResidualVQ  # type: ignore

    embeds, x = generic_rvq_inputs(seed, n_q, size, batch, frames)
    out = {"seed": np.int64(seed), "n_q": np.int64(n_q), "size": np.int64(size), "batch": np.int64(batch),
           "frames": np.int64(frames)}
    for channel_last in (False, True):
        vq = ResidualVQ(n_q, dim=128, codebook_size=size, channel_last=channel_last).eval()
        assert [k for k in vq.state_dict()][:4] == ["layers.0._codebook.initted", "layers.0._codebook.embed",
                                                    "layers.0._codebook.ema_embed", "layers.0._codebook.ema_num"]
        for layer, e in zip(vq.layers, embeds):
            layer._codebook.embed.copy_(torch.from_numpy(e))
        xin = torch.from_numpy(x if not channel_last else np.ascontiguousarray(x.transpose(0, 2, 1)))
        for n in (None, 3, 1):
            with torch.no_grad():
                q, num_replaces, loss = vq(xin, n)
            assert num_replaces.dtype == np.int64 and num_replaces.shape == (n_q,) and not num_replaces.any()
            tag = f"{'cl' if channel_last else 'cf'}_{n}"
            out[f"q_{tag}"] = q.numpy()
            out[f"loss_{tag}"] = np.float32(loss.item())
            print(name, tag, tuple(q.shape), float(loss))
    np.savez_compressed(os.path.join(HERE, name), **out)


if __name__ == "__main__":
    torch.set_num_threads(8)
    if sys.argv[1:] == ["generic_rvq"]:
        ref_generic_rvq("ref_generic_rvq.npz")
        sys.exit(0)
    if sys.argv[1:] == ["music"]:
        ref_music_published("ref_music_published.npz")
        sys.exit(0)
    if sys.argv[1:] == ["train"]:
        ref_train("ref_train_random.npz", 6, 3, 5, batch=2, lengths=[320 * 8, 320 * 12 + 77, 333, 1])
        sys.exit(0)
    speech_kat()
    ref_random("ref_random_speech.npz", 8, 1, batch=2, frames=12, stream_hops=1)
    ref_random("ref_random_music.npz", 12, 2, batch=3, frames=10, stream_hops=3)
    ref_train("ref_train_random.npz", 6, 3, 5, batch=2, lengths=[320 * 8, 320 * 12 + 77, 333, 1])
    ref_music_published("ref_music_published.npz")
    ref_generic_rvq("ref_generic_rvq.npz")
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))
