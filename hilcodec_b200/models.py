"""Training-graph call signatures on top of the deployment kernels.

The reference's validation / PESQ / inference loops (`models/hilcodec/wrapper.py:347, 373, 391`) call the
training-graph `HILCodec.forward(x, n) -> (wav, num_replaces, loss_vq)` (`models/hilcodec/models.py:111-118`)
and `ResidualVQ.forward(x[B,C,T], n, return_indices)` (`models/hilcodec/vector_quantize.py:199-243`).
These adapters expose those signatures and shapes over the CUDA path.

Caveat, stated rather than hidden: the arithmetic is the DEPLOYMENT graph's (`streaming.py`).  The
reference's own two graphs differ in the decoder (`pre_scale`, `conv_post` bias scaling -- SURVEY.md
quirks 1-2) and in the VQ distance formula (quirk 4); encoders agree to ~4e-6 and give identical
indices.  Inputs must be a multiple of the hop length (the training graph's right "extra padding",
`modules/conv.py:61-68`, is not reproduced).
"""
from __future__ import annotations

import typing as tp

import numpy as np
import torch
from torch import Tensor, nn

from . import streaming


class ResidualVQ(nn.Module):
    """`vector_quantize.py:179` eval-mode signature: x [B,C,T] -> (quantized [B,C,T], num_replaces, loss[, indices [B,n,T]])."""

    def __init__(self, deploy: streaming.ResidualVQ):
        super().__init__()
        self._vq = deploy
        self.num_quantizers = len(deploy.layers)

    def forward(self, x: Tensor, n: tp.Optional[int] = None, return_indices: bool = False):
        n = self.num_quantizers if n is None else n
        assert 1 <= n <= self.num_quantizers, f"n={n} out of range (vector_quantize.py:213)"
        xt = x.transpose(1, 2).contiguous()                       # [B,T,C], vector_quantize.py:206
        idx, q = self._vq.quantize(xt, n)
        quantized_out = q.transpose(1, 2)
        loss = torch.nn.functional.mse_loss(x, quantized_out)      # vector_quantize.py:235
        num_replaces = np.zeros(self.num_quantizers, dtype=np.int64)  # eval: no dead-code replacement
        if return_indices:
            return quantized_out, num_replaces, loss, idx.permute(1, 0, 2)  # [B,n,T]
        return quantized_out, num_replaces, loss


class HILCodec(nn.Module):
    """`models/hilcodec/models.py:24` call contract: forward(x[B,1,T], n=None) -> (wav.float(), num_replaces, loss_vq)."""

    def __init__(self, deploy: streaming.HILCodec):
        super().__init__()
        self.deploy = deploy
        self.quantizer = ResidualVQ(deploy.quantizer)
        self.sample_rate = deploy.sample_rate
        self.channels = deploy.channels

    @classmethod
    def from_pretrained(cls, name: str) -> "HILCodec":
        return cls(streaming.HILCodec.from_pretrained(name))

    def encode(self, x: Tensor) -> Tensor:
        """SEANetEncoder call shape: [B,1,T] -> [B,C,T/hop] (channel-first)."""
        ce = self.deploy.encoder.initialize_cache(x)
        z, _ = self.deploy.encoder(x, *ce)
        return z.transpose(1, 2)

    def decode(self, q: Tensor) -> Tensor:
        """SEANetDecoder call shape: [B,C,F] -> [B,1,hop*F]."""
        cd = self.deploy.decoder.initialize_cache(q)
        y, _ = self.deploy.decoder(q.transpose(1, 2).contiguous(), *cd)
        return y

    @torch.no_grad()
    def forward(self, x: Tensor, n: tp.Optional[int] = None):
        z = self.encode(x)
        q, num_replaces, loss_vq = self.quantizer(z, n)
        return self.decode(q).float(), num_replaces, loss_vq

    def remove_weight_reparameterizations(self) -> None:
        return None
