"""Bitstream packing of RVQ indices: 10 bits per codebook per frame.

The reference stores indices as int16 `.npy` (`test_onnx.py:99`); on a wire HILCodec's nominal
rate is 0.75 kbps per codebook.  Packing runs on the GPU (`csrc/bitpack.cu`); `pack_numpy` /
`unpack_numpy` state the format for host-side consumers and tests.

Format: frame-major, `bytes_per_frame = ceil(n * bits / 8)`; inside a frame the n indices are
concatenated LSB first, `bits = log2(codebook_size)` each.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


def bytes_per_frame(n: int, bits: int = 10) -> int:
    return (n * bits + 7) // 8


def pack(model, indices: torch.Tensor) -> torch.Tensor:
    """indices [n,B,F] int64 (CUDA) -> uint8 [B,F,bytes_per_frame] (CUDA)."""
    if not indices.is_cuda or indices.dtype != torch.int64 or indices.dim() != 3:
        raise TypeError("indices must be a CUDA int64 tensor [n, B, F]")
    n, B, F = indices.shape
    core, dev = model._core, indices.device
    lib = _lib.load()
    with torch.cuda.device(dev):
        h = core.model(dev)
        bpf = lib.hil_bitstream_bytes_per_frame(h, n)
        if bpf <= 0:
            raise ValueError("n must satisfy 1 <= n <= num_quantizers")
        out = torch.empty(B, F, bpf, dtype=torch.uint8, device=dev)
        _lib.check(lib.hil_pack_indices(h, indices.contiguous().data_ptr(), B, F, n, out.data_ptr(),
                                        torch.cuda.current_stream(dev).cuda_stream))
    return out


def unpack(model, packed: torch.Tensor, n: int) -> torch.Tensor:
    """uint8 [B,F,bytes_per_frame] (CUDA) -> indices [n,B,F] int64 (CUDA)."""
    if not packed.is_cuda or packed.dtype != torch.uint8 or packed.dim() != 3:
        raise TypeError("packed must be a CUDA uint8 tensor [B, F, bytes_per_frame]")
    B, F, bpf = packed.shape
    core, dev = model._core, packed.device
    lib = _lib.load()
    with torch.cuda.device(dev):
        h = core.model(dev)
        if lib.hil_bitstream_bytes_per_frame(h, n) != bpf:
            raise ValueError(f"bytes per frame {bpf} does not match n={n}")
        idx = torch.empty(n, B, F, dtype=torch.int64, device=dev)
        _lib.check(lib.hil_unpack_indices(h, packed.contiguous().data_ptr(), B, F, n, idx.data_ptr(),
                                          torch.cuda.current_stream(dev).cuda_stream))
    return idx


def pack_numpy(indices: np.ndarray, bits: int = 10) -> np.ndarray:
    """Host statement of the format: indices [n,B,F] -> uint8 [B,F,bytes_per_frame]."""
    n, B, F = indices.shape
    bpf = bytes_per_frame(n, bits)
    out = np.zeros((B, F, bpf), dtype=np.uint8)
    for b in range(B):
        for f in range(F):
            acc = 0
            for s in range(n):
                acc |= (int(indices[s, b, f]) & ((1 << bits) - 1)) << (s * bits)
            out[b, f] = np.frombuffer(acc.to_bytes(bpf, "little"), dtype=np.uint8)
    return out


def unpack_numpy(packed: np.ndarray, n: int, bits: int = 10) -> np.ndarray:
    B, F, bpf = packed.shape
    out = np.zeros((n, B, F), dtype=np.int64)
    for b in range(B):
        for f in range(F):
            acc = int.from_bytes(packed[b, f].tobytes(), "little")
            for s in range(n):
                out[s, b, f] = (acc >> (s * bits)) & ((1 << bits) - 1)
    return out
