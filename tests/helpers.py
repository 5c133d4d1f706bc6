"""Shared helpers of the parity tests."""
import os

import numpy as np
import torch

from hilcodec_b200 import weights as W
from oracle import hilcodec_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def synth_wav(batch, samples, seed=1234):
    g = torch.Generator().manual_seed(seed)
    return (0.1 * torch.randn(batch, 1, samples, generator=g)).clamp(-1, 1)


def params(w):
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in w.items()}


def oracle_cfg(n_q):
    return O.CodecConfig(num_quantizers=n_q)


def index_report(cfg, p, z_gpu, idx_gpu, idx_ref, n):
    """Near-tie policy: every decision where the CUDA path and the CPU
    reference disagree must be a near tie.  A disagreement at stage s desynchronises the
    residual of that frame, so only the FIRST differing stage of a frame is judged; its
    relative gap (d2-d1)/d1 is measured in float64 along the GPU path's own residual.
    Returns (n_mismatched_frames, worst_gap)."""
    idx_gpu = idx_gpu.cpu()
    diff = (idx_gpu != idx_ref)
    if not diff.any():
        return 0, 0.0
    _, gaps = O.rvq_margins(cfg, p, z_gpu.cpu(), n)
    # first differing stage per frame
    first = torch.argmax(diff.int(), dim=0)
    frames = diff.any(dim=0)
    worst = 0.0
    for b, f in zip(*torch.nonzero(frames, as_tuple=True)):
        worst = max(worst, float(gaps[first[b, f], b, f]))
    return int(frames.sum()), worst
