"""Host restatement of the index bitstream format packed by `hilcodec_b200/csrc/bitpack.cu`.

TEST INFRASTRUCTURE ONLY (never imported by the product package).  The reference has no bitstream: it stores
indices as an int16 `.npy` (`test_onnx.py:99`); the format is this repo's and is pinned by
the known-answer bytes in tests/test_bitstream_cpu.py.

Format: frame-major, `bytes_per_frame = ceil(n * bits / 8)`; inside a frame the n indices are concatenated
LSB first, `bits = log2(codebook_size)` each (10 bits: 0.75 kbps per codebook at 75 frames/s).
"""
from __future__ import annotations

import numpy as np


def bytes_per_frame(n: int, bits: int = 10) -> int:
    return (n * bits + 7) // 8


def pack_numpy(indices: np.ndarray, bits: int = 10) -> np.ndarray:
    """indices [n,B,F] -> uint8 [B,F,bytes_per_frame]."""
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
