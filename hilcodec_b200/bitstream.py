"""Bitstream packing of RVQ indices: 10 bits per codebook per frame.

The reference stores indices as int16 `.npy` (`test_onnx.py:99`); on a wire HILCodec's nominal
rate is 0.75 kbps per codebook.  Packing runs on the GPU (`csrc/bitpack.cu`) and only there; the host-side
restatement of the format used by the tests is `oracle/bitstream_oracle.py`.

Format: frame-major, `bytes_per_frame = ceil(n * bits / 8)`; inside a frame the n indices are
concatenated LSB first, `bits = log2(codebook_size)` each.
"""
from __future__ import annotations

import ctypes as C

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
