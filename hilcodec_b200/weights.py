"""Weight containers for the deployment graph: config, tensor names/shapes,
DFT bases, seeded random initialisation and .npz loading.

Tensor names are the `state_dict()` keys of the reference's
`models/hilcodec/streaming.py` modules after folding, prefixed
with `encoder.` / `decoder.`, plus `quantizer.layers.{i}.embed`.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import Dict, List, Tuple

import numpy as np

WAV_STD = 0.1122080159  # streaming.py:382 / :532


@dataclass(frozen=True)
class CodecConfig:
    """Architecture constants (configs/hilcodec_{speech,music}.yaml:2-38)."""
    channels_enc: int = 64
    channels_dec: int = 96
    n_fft_base: int = 64
    n_residual_enc: int = 2
    n_residual_dec: int = 3
    res_scale_enc: float = 0.5773502691896258
    res_scale_dec: float = 0.5773502691896258
    strides: Tuple[int, ...] = (8, 5, 4, 2)
    kernel_size: int = 5
    dim: int = 128
    codebook_size: int = 1024
    num_quantizers: int = 8

    @property
    def hop(self) -> int:
        return int(math.prod(self.strides))

    @property
    def enc_ratios(self) -> Tuple[int, ...]:
        return tuple(reversed(self.strides))


HIL_SPEECH = CodecConfig(num_quantizers=8)
HIL_MUSIC = CodecConfig(num_quantizers=12)
CONFIGS = {"hil_speech": HIL_SPEECH, "hil_music": HIL_MUSIC}


def dft_basis(n_fft: int) -> np.ndarray:
    """[2F,1,N] rows [cos; sin] * periodic hann, norm "backward"
    (causal_layers.py:115-129).  Built with torch float32 ops so it is bit-identical
    to the buffer the reference registers."""
    import torch

    n = torch.arange(n_fft, dtype=torch.float32).view(1, 1, n_fft)
    k = torch.arange(n_fft // 2 + 1, dtype=torch.float32).view(-1, 1, 1)
    window = torch.hann_window(n_fft)
    cos = torch.cos(-2 * math.pi / n_fft * k * n)
    sin = torch.sin(-2 * math.pi / n_fft * k * n)
    return (torch.cat([cos, sin], dim=0) * window).numpy()


def tensor_shapes(cfg: CodecConfig) -> Dict[str, Tuple[int, ...]]:
    """Every tensor the deployment graph consumes, in streaming.py's naming."""
    k = cfg.kernel_size
    s: Dict[str, Tuple[int, ...]] = {}
    e = "encoder."
    c = cfg.channels_enc
    s[e + "conv_pre.weight"] = (c, 1, k)
    s[e + "conv_pre.bias"] = (c,)
    for i, r in enumerate(cfg.enc_ratios):
        n_fft = cfg.n_fft_base * 2 ** i
        f = n_fft // 2 + 1
        s[f"{e}spec_blocks.{i}.spec.weight"] = (2 * f, 1, n_fft)
        s[f"{e}spec_blocks.{i}.layer.weight"] = (c, f, 1)
        s[f"{e}spec_blocks.{i}.layer.bias"] = (c,)
        for j in range(cfg.n_residual_enc):
            for b in range(2):
                pre = f"{e}blocks.{i}.{j}.block.{b}."
                s[pre + "pointwise.1.weight"] = (c, c, 1)
                s[pre + "depthwise.weight"] = (c, 1, k)
                s[pre + "depthwise.bias"] = (c,)
        s[f"{e}downsample_pointwise.{i}.1.weight"] = (2 * c, c, 1)
        s[f"{e}downsample_depthwise.{i}.weight"] = (2 * c, 1, 2 * r)
        s[f"{e}downsample_depthwise.{i}.bias"] = (2 * c,)
        c *= 2
    n_fft = cfg.n_fft_base * 2 ** len(cfg.strides)
    f = n_fft // 2 + 1
    s[e + "spec_post.spec.weight"] = (2 * f, 1, n_fft)
    s[e + "spec_post.layer.weight"] = (c, f, 1)
    s[e + "spec_post.layer.bias"] = (c,)
    s[e + "conv_post_depthwise.weight"] = (c, 1, k)
    s[e + "conv_post_pointwise.weight"] = (cfg.dim, c, 1)
    s[e + "conv_post_pointwise.bias"] = (cfg.dim,)

    d = "decoder."
    c = cfg.channels_dec * 2 ** len(cfg.strides)
    s[d + "conv_pre_pointwise.weight"] = (c, cfg.dim, 1)
    s[d + "conv_pre_depthwise.weight"] = (c, 1, k)
    s[d + "conv_pre_depthwise.bias"] = (c,)
    for i, r in enumerate(cfg.strides):
        s[f"{d}upsample_depthwise.{i}.weight"] = (c, 1, 2 * r)
        s[f"{d}upsample_pointwise.{i}.weight"] = (c // 2, c, 1)
        s[f"{d}upsample_pointwise.{i}.bias"] = (c // 2,)
        for j in range(cfg.n_residual_dec):
            for b in range(2):
                pre = f"{d}blocks.{i}.{j}.block.{b}."
                s[pre + "pointwise.1.weight"] = (c // 2, c // 2, 1)
                s[pre + "depthwise.weight"] = (c // 2, 1, k)
                s[pre + "depthwise.bias"] = (c // 2,)
        c //= 2
    s[d + "conv_post.weight"] = (1, c, k)
    s[d + "conv_post.bias"] = (1,)
    for i in range(cfg.num_quantizers):
        s[f"quantizer.layers.{i}.embed"] = (cfg.codebook_size, cfg.dim)
    return s


def random_weights(cfg: CodecConfig, seed: int = 0) -> Dict[str, np.ndarray]:
    """Seeded random folded weights of the right shapes (numpy PCG64, float32).

    Scaled so activations keep O(1) magnitude through both stacks (variance-preserving
    pointwise convs, ~unit-gain depthwise taps, residual branches at 1/sqrt(3) like the
    trained `res_scale`), which keeps the log-STFT branch, ELU's negative side, tanh and
    the codebook search all exercised by parity tests without pretrained weights.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    out: Dict[str, np.ndarray] = {}
    for name, shape in tensor_shapes(cfg).items():
        if name.endswith("spec.weight"):
            out[name] = dft_basis(shape[2])
            continue
        if name.endswith(".embed"):
            stage = int(name.split(".")[2])
            v = rng.standard_normal(shape, dtype=np.float32) * np.float32(0.7 ** stage)
        elif name.endswith("bias"):
            v = rng.standard_normal(shape, dtype=np.float32) * np.float32(0.05)
        elif "conv_pre.weight" in name:
            v = rng.standard_normal(shape, dtype=np.float32) * np.float32(0.45 / WAV_STD)
        elif "spec" in name and name.endswith("layer.weight"):
            v = rng.standard_normal(shape, dtype=np.float32) * np.float32(0.1 / math.sqrt(shape[1]))
        elif "depthwise" in name:
            taps = shape[2]
            v = rng.standard_normal(shape, dtype=np.float32) * np.float32(1.0 / math.sqrt(taps))
            if ".block.1." in name:
                v *= np.float32(0.5773502691896258)
            if "upsample" in name:
                v *= np.float32(math.sqrt(2.0))  # each output sees 2 of the 2r taps
        elif name == "decoder.conv_post.weight":
            v = rng.standard_normal(shape, dtype=np.float32) * np.float32(0.5 / math.sqrt(shape[1] * shape[2]))
        else:  # pointwise 1x1
            v = rng.standard_normal(shape, dtype=np.float32) * np.float32(1.2 / math.sqrt(shape[1]))
        out[name] = v.astype(np.float32)
    # residual-branch bias follows the folded scale too
    for name in list(out):
        if ".block.1.depthwise.bias" in name:
            out[name] = out[name] * np.float32(0.5773502691896258)
    return out


def weights_dir() -> str:
    return os.environ.get(
        "HILCODEC_WEIGHTS", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "weights"))


def pretrained_path(name: str) -> str:
    return os.path.join(weights_dir(), f"{name}.npz")


def have_pretrained(name: str) -> bool:
    return os.path.exists(pretrained_path(name))


def load_npz(path: str) -> Dict[str, np.ndarray]:
    with np.load(path) as d:
        return {k: np.ascontiguousarray(d[k], dtype=np.float32) for k in d.files}


def load_pretrained(name: str) -> Dict[str, np.ndarray]:
    """Published weights extracted by `python -m hilcodec_b200.onnx_weights`."""
    path = pretrained_path(name)
    if not os.path.exists(path):
        raise FileNotFoundError(
            f"{path} missing: run `python -m hilcodec_b200.onnx_weights <reference>/onnx {name} {path}`")
    return load_npz(path)


def check_weights(cfg: CodecConfig, w: Dict[str, np.ndarray]) -> None:
    """Raise on a missing tensor or wrong shape (unexpected extras are ignored like
    `load_state_dict(strict=False)` ignores `ema_num`)."""
    for name, shape in tensor_shapes(cfg).items():
        if name not in w:
            raise KeyError(f"missing tensor {name}")
        if tuple(w[name].shape) != shape:
            raise ValueError(f"{name}: shape {tuple(w[name].shape)} != {shape}")


def cache_shapes(cfg: CodecConfig, batch: int) -> Tuple[List[Tuple[int, int, int]], List[Tuple[int, int, int]]]:
    """(encoder 22, decoder 30) cache shapes in the reference's list order
    (streaming.py:458-470, :599-607)."""
    k = cfg.kernel_size
    n_fft_post = cfg.n_fft_base * 2 ** len(cfg.strides)
    enc = [(batch, 1, n_fft_post - 1)]
    c = cfg.channels_enc
    for r in cfg.enc_ratios:
        enc += [(batch, c, k - 1)] * (2 * cfg.n_residual_enc)
        enc.append((batch, 2 * c, r))  # d(k-1)-(s-1) with k=2r, s=r
        c *= 2
    enc.append((batch, c, k - 1))
    c = cfg.channels_dec * 2 ** len(cfg.strides)
    dec = [(batch, c, k - 1)]
    for r in cfg.strides:
        dec.append((batch, c, (2 * r - 1) // r))
        dec += [(batch, c // 2, k - 1)] * (2 * cfg.n_residual_dec)
        c //= 2
    dec.append((batch, c, k - 1))
    return enc, dec
