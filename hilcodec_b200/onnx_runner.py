"""`test_onnx.py`-compatible runner backed by the B200 kernels.

Mirrors the reference runner's workflow and artefacts (`test_onnx.py:50-139, 142-189`):

    python -m hilcodec_b200.onnx_runner -n hil_speech -q 8 --enc --dec [-f 1] [-H 320] \\
        [--input onnx/input_speech.wav] [--outdir onnx]

* encoder pass: the clip is cut to a multiple of the hop, fed hop by hop (`-H`, 320) through
  `Encoder` + `ResidualVQ` with the caches handed back and forth exactly like `e_in{i}`/`e_out{i}`,
  and the indices are written as int16 `[n, B, T]` to `{outdir}/{name}_quantized.npy`
  (`test_onnx.py:95-100`);
* decoder pass: reads that file, dequantises and decodes `-f` frames per call with the `d_in{i}`/
  `d_out{i}` cache protocol and writes `{outdir}/{name}_output.wav` (`test_onnx.py:103-139`);
* prints the reference's timer lines: `encoder: .. s / rtf: .. (up)`.
The `.onnx` graphs themselves are not executed; weights come from `weights/{name}.npz`
(extracted from those graphs by `hilcodec_b200.onnx_weights`).
"""
from __future__ import annotations

import argparse
import os
import time

import numpy as np
import torch


class Timer:
    """test_onnx.py:20-47"""

    def __init__(self, sr: int):
        self.sr = sr
        self.enc_time = 0.0
        self.dec_time = 0.0
        self.start_time = time.perf_counter()
        self.wav_len = 0

    def tic(self):
        torch.cuda.synchronize()
        self.start_time = time.perf_counter()

    def encoder_time(self):
        torch.cuda.synchronize()
        et = time.perf_counter()
        self.enc_time += et - self.start_time
        self.start_time = et

    def decoder_time(self):
        torch.cuda.synchronize()
        et = time.perf_counter()
        self.dec_time += et - self.start_time
        self.start_time = et

    def print(self):
        wav_time = self.wav_len / self.sr
        print(f"\rwav length: {wav_time:.1f} s")
        if self.enc_time > 0:
            print(f"encoder: {self.enc_time:.1f} s / rtf: {wav_time/self.enc_time:.4f} (↑)")
        if self.dec_time > 0:
            print(f"decoder: {self.dec_time:.1f} s / rtf: {wav_time/self.dec_time:.4f} (↑)")


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


def encoder(model, wav: np.ndarray, hop_size: int, num_quantizers: int, timer: Timer) -> np.ndarray:
    length = len(wav) // hop_size * hop_size
    x = torch.from_numpy(wav[:length]).view(1, 1, -1).cuda()
    timer.wav_len = length
    cache = model.encoder.initialize_cache(x)
    indices = []
    timer.tic()
    for i in range(0, length, hop_size):
        z, cache = model.encoder(x[:, :, i:i + hop_size], *cache)
        indices.append(model.quantizer(z, num_quantizers))
    timer.encoder_time()
    return torch.cat(indices, dim=2).cpu().numpy().astype(np.int16)  # [n, B, T]


def decoder(model, indices: np.ndarray, num_frames: int, num_quantizers: int, timer: Timer) -> np.ndarray:
    idx = torch.from_numpy(indices.astype(np.int64)).cuda()
    cache = model.decoder.initialize_cache(torch.zeros(idx.shape[1], 1, 1, device="cuda"))
    out = []
    timer.tic()
    for i in range(0, idx.shape[2], num_frames):
        q = model.dequantizer(idx[:, :, i:i + num_frames], num_quantizers)
        y, cache = model.decoder(q, *cache)
        out.append(y)
    timer.decoder_time()
    wav_out = torch.cat(out, dim=2)[0, 0].cpu().numpy()
    timer.wav_len = len(wav_out)
    return wav_out


def main(argv=None) -> int:
    from .streaming import HILCodec

    ap = argparse.ArgumentParser(description=__doc__.splitlines()[0])
    ap.add_argument("-n", "--name", default="hil_speech")
    ap.add_argument("-q", "--num_quantizers", type=int, default=8)
    ap.add_argument("-f", "--num_frames", type=int, default=1)
    ap.add_argument("-H", "--hop_size", type=int, default=320)
    ap.add_argument("--sr", type=int, default=24000)
    ap.add_argument("--enc", action="store_true")
    ap.add_argument("--dec", action="store_true")
    ap.add_argument("--input", default="onnx/input_speech.wav")
    ap.add_argument("--outdir", default="onnx")
    a = ap.parse_args(argv)
    model = HILCodec.from_pretrained(a.name).cuda()
    os.makedirs(a.outdir, exist_ok=True)
    timer = Timer(a.sr)
    qpath = os.path.join(a.outdir, f"{a.name}_quantized.npy")
    if a.enc:
        np.save(qpath, encoder(model, load_wav(a.input, a.sr), a.hop_size, a.num_quantizers, timer))
    if a.dec:
        from scipy.io import wavfile

        wav_out = decoder(model, np.load(qpath), a.num_frames, a.num_quantizers, timer)
        wavfile.write(os.path.join(a.outdir, f"{a.name}_output.wav"), a.sr,
                      np.clip(np.round(wav_out * 32768.0), -32768, 32767).astype(np.int16))
    timer.print()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
