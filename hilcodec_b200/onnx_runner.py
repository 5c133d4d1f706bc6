"""`test_onnx.py`-compatible runner backed by the B200 kernels.

Same workflow, flags, tensor names and artefacts as the reference's ONNXRuntime runner (`test_onnx.py:50-139`,
flags `:169-189`), with the `.onnx` sessions replaced by the drop-in modules:

    python -m hilcodec_b200.onnx_runner -n hil_speech -q 8 --enc --dec [-f 1] [-H 320] [-t 1] \\
        [--input onnx/input_speech.wav] [--outdir onnx] [--cache-dir onnx]

* encoder pass (`test_onnx.py:50-100`): the clip is cut to a multiple of `hop_size * num_frames` -- the reference
  multiplies the hop by `-f` for BOTH passes (`:153`) -- and fed chunk by chunk as `wav_in`; the 22 caches travel in a
  dict under the graph's input names `e_in0 .. e_in21`, initialised from `{cache_dir}/{name}_cache_enc.npz` when that
  file exists (the exporter writes all-zero caches there, `:73`) and refreshed from the outputs `e_out{i}` after every
  call (`:80-81`); the indices are written as int16 `[n, B, T]` to `{outdir}/{name}_quantized.npy` (`:95-100`);
* decoder pass (`:103-139`): reads that file, dequantises `-f` frames per call (`q`), decodes with the `d_in{i}` /
  `d_out{i}` dict (`{name}_cache_dec.npz`, `:121, 133-134`) and writes `{outdir}/{name}_output.wav`;
* prints the reference's timing lines (`:41-47`): `wav length`, `encoder: .. s / rtf: .. (up)`, `decoder: ..`.
`-t / --num_threads` is accepted for command-line compatibility; it pins the HOST threads (torch / BLAS), the kernels
run on the GPU either way.  `--save-cache` writes the two zero-cache `.npz` files in the reference's naming.
The `.onnx` graphs themselves are not executed; weights come from `weights/{name}.npz` (extracted from those graphs by
`hilcodec_b200.onnx_weights`).
"""
from __future__ import annotations

import argparse
import os
import time
import typing as tp

import numpy as np
import torch


def load_wav(path: str, sr: int) -> np.ndarray:
    """What `librosa.load(path, sr=sr)` returns for a mono PCM16 file already at `sr`."""
    from scipy.io import wavfile

    file_sr, wav = wavfile.read(path)
    if file_sr != sr:
        raise ValueError(f"{path}: sample rate {file_sr} != {sr} (resampling is out of scope here)")
    if wav.ndim > 1:
        wav = wav.mean(axis=1)
    if wav.dtype == np.int16:
        wav = wav.astype(np.float32) / 32768.0
    return wav.astype(np.float32)


def _elapsed(fn: tp.Callable[[], tp.Any]) -> tp.Tuple[tp.Any, float]:
    """Wall-clock seconds of `fn()` with the device drained on both sides."""
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return out, time.perf_counter() - t0


def report(samples: int, sr: int, enc_s: float, dec_s: float) -> str:
    """The lines `test_onnx.py` prints at the end of a run (`:41-47`); rtf = audio seconds / compute seconds."""
    audio_s = samples / sr
    lines = [f"\rwav length: {audio_s:.1f} s"]
    for label, sec in (("encoder", enc_s), ("decoder", dec_s)):
        if sec > 0:
            lines.append(f"{label}: {sec:.1f} s / rtf: {audio_s / sec:.4f} (↑)")
    return "\n".join(lines)


# ------------------------------------------------------------------------------------------ cache dict protocol
def cache_names(side: str, count: int) -> tp.List[str]:
    """Graph input names of the caches: `e_in{i}` (encoder, 22) / `d_in{i}` (decoder, 30)."""
    return [f"{'e' if side == 'enc' else 'd'}_in{i}" for i in range(count)]


def load_cache_dict(path: tp.Optional[str], side: str, zero: tp.Sequence[torch.Tensor]) -> tp.Dict[str, torch.Tensor]:
    """`dict(np.load("onnx/{name}_cache_{enc,dec}.npz"))` (`test_onnx.py:73, 121`) as CUDA tensors; zero caches of the
    right shapes when the file is absent.  Shapes are checked against the model's."""
    names = cache_names(side, len(zero))
    if path is None or not os.path.isfile(path):
        return dict(zip(names, zero))
    npz = np.load(path)
    missing = [k for k in names if k not in npz.files]
    if missing:
        raise KeyError(f"{path}: missing cache tensors {missing[:4]}")
    out = {}
    for k, z in zip(names, zero):
        a = npz[k]
        if tuple(a.shape) != tuple(z.shape):
            raise ValueError(f"{path}: {k} has shape {a.shape}, the model needs {tuple(z.shape)}")
        out[k] = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(z.device)
    return out


def save_zero_caches(model, outdir: str, name: str, batch: int = 1) -> None:
    probe = torch.zeros(batch, 1, 1, device="cuda")
    for side, mod in (("enc", model.encoder), ("dec", model.decoder)):
        zero = mod.initialize_cache(probe)
        np.savez(os.path.join(outdir, f"{name}_cache_{side}.npz"),
                 **{k: z.cpu().numpy() for k, z in zip(cache_names(side, len(zero)), zero)})


# ------------------------------------------------------------------------------------------ the two passes
def encoder(model, wav: np.ndarray, hop_size: int, num_quantizers: int, cache_path: tp.Optional[str] = None):
    """-> (indices int16 [n, B, T], samples consumed, seconds)."""
    length = len(wav) // hop_size * hop_size
    x = torch.from_numpy(wav[:length]).view(1, 1, -1).cuda()
    enc_input = load_cache_dict(cache_path, "enc", model.encoder.initialize_cache(x))
    names = cache_names("enc", len(enc_input))

    def run():
        chunks = []
        for i in range(0, length, hop_size):
            enc_input["wav_in"] = x[:, :, i:i + hop_size]
            z, out = model.encoder(enc_input["wav_in"], *[enc_input[k] for k in names])
            for k, c in zip(names, out):          # e_in{i} <- e_out{i}
                enc_input[k] = c
            chunks.append(model.quantizer(z, num_quantizers))
        return torch.cat(chunks, dim=2)

    indices, sec = _elapsed(run)
    return indices.cpu().numpy().astype(np.int16), length, sec


def decoder(model, indices: np.ndarray, num_frames: int, num_quantizers: int, cache_path: tp.Optional[str] = None):
    """-> (wav float32 [T], seconds)."""
    idx = torch.from_numpy(indices.astype(np.int64)).cuda()
    dec_input = load_cache_dict(cache_path, "dec", model.decoder.initialize_cache(torch.zeros(idx.shape[1], 1, 1, device="cuda")))
    names = cache_names("dec", len(dec_input))

    def run():
        pieces = []
        for i in range(0, idx.shape[2], num_frames):
            dec_input["q"] = model.dequantizer(idx[:, :, i:i + num_frames], num_quantizers)
            y, out = model.decoder(dec_input["q"], *[dec_input[k] for k in names])
            for k, c in zip(names, out):          # d_in{j} <- d_out{j}
                dec_input[k] = c
            pieces.append(y)
        return torch.cat(pieces, dim=2)

    wav_out, sec = _elapsed(run)
    return wav_out[0, 0].cpu().numpy(), sec


def main(argv=None) -> int:
    from .streaming import HILCodec

    ap = argparse.ArgumentParser(description=__doc__.splitlines()[0])
    ap.add_argument("-n", "--name", type=str, default="hil_speech", help="Model name. Default: hil_speech")
    ap.add_argument("-q", "--num_quantizers", type=int, default=8, help="Number of quantizers to use. Default: 8")
    ap.add_argument("-t", "--num_threads", type=int, default=1, help="Number of (host) threads to use. Default: 1")
    ap.add_argument("-f", "--num_frames", type=int, default=1, help="Number of frames to process at once. Default: 1")
    ap.add_argument("-H", "--hop_size", type=int, default=320, help="Hop size. Default: 320")
    ap.add_argument("--enc", action="store_true", help="Run encoder")
    ap.add_argument("--dec", action="store_true", help="Run decoder")
    ap.add_argument("--sr", type=int, default=24_000, help="Sampling rate. Default: 24000")
    ap.add_argument("--input", default="onnx/input_speech.wav")
    ap.add_argument("--outdir", default="onnx")
    ap.add_argument("--cache-dir", default=None, help="where {name}_cache_{enc,dec}.npz live (default: --outdir)")
    ap.add_argument("--save-cache", action="store_true", help="write the zero-cache .npz files and exit")
    a = ap.parse_args(argv)
    torch.set_num_threads(max(1, a.num_threads))
    model = HILCodec.from_pretrained(a.name).cuda()
    os.makedirs(a.outdir, exist_ok=True)
    cache_dir = a.cache_dir or a.outdir
    if a.save_cache:
        save_zero_caches(model, cache_dir, a.name)
        return 0
    hop = a.hop_size * a.num_frames                     # test_onnx.py:153
    qpath = os.path.join(a.outdir, f"{a.name}_quantized.npy")
    samples, enc_s, dec_s = 0, 0.0, 0.0
    if a.enc:
        indices, samples, enc_s = encoder(model, load_wav(a.input, a.sr), hop, a.num_quantizers,
                                          os.path.join(cache_dir, f"{a.name}_cache_enc.npz"))
        np.save(qpath, indices)
    if a.dec:
        from scipy.io import wavfile

        wav_out, dec_s = decoder(model, np.load(qpath), a.num_frames, a.num_quantizers,
                                 os.path.join(cache_dir, f"{a.name}_cache_dec.npz"))
        samples = len(wav_out)
        wavfile.write(os.path.join(a.outdir, f"{a.name}_output.wav"), a.sr,
                      np.clip(np.round(wav_out * 32768.0), -32768, 32767).astype(np.int16))
    print(report(samples, a.sr, enc_s, dec_s))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
