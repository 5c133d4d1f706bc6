"""The reference's TRAINING-graph interface on top of the CUDA path.

The reference's validation / PESQ / inference loops (`models/hilcodec/wrapper.py:347, 373, 391`) call the
training-graph `HILCodec.forward(x, n) -> (wav, num_replaces, loss_vq)` (`models/hilcodec/models.py:111-118`),
whose sub-modules are `SEANetEncoder(x[B,1,T]) -> [B,128,ceil(T/hop)]` (`modules/seanet.py:368-378`),
`ResidualVQ(x[B,C,T], n, return_indices)` (`models/hilcodec/vector_quantize.py:199-243`) and
`SEANetDecoder(z[B,128,F]) -> [B,1,hop*F]` (`modules/seanet.py:477-479`).  This module exposes those
signatures in eval mode (no EMA update, no code expiry, no straight-through term).

Two arithmetic variants, chosen with `graph`:

* `graph="train"` -- the training graph itself (`hil_model_set_graph(HIL_GRAPH_TRAIN)`): decoder ResBlock j
  scaled by (1 + j/3)^-0.5, `Scale(wav_std)` after conv_post's bias, codebook search without the sum(x^2)
  term, and any input length (each causal conv pads its own right edge, `modules/conv.py:61-68, 222-236`;
  `hil_encode_ragged`).  This is what a checkpoint scores in the reference's own validation loop.
* `graph="deploy"` -- the deployment graph's arithmetic (`streaming.py`; the one the ONNX export and the golden
  vectors use) behind the same signatures; inputs must be a multiple of the hop length.

The reference's two graphs are NOT numerically equal in the decoder: with the published
weights the decoded PCM differs at the 1e-2 level, the encoder and the indices agree.
"""
from __future__ import annotations

import typing as tp

import numpy as np
import torch
from torch import Tensor, nn

from . import _lib, checkpoint, fold, streaming
from .weights import CONFIGS, CodecConfig, load_pretrained

_GRAPHS = {"deploy": _lib.HIL_GRAPH_DEPLOY, "train": _lib.HIL_GRAPH_TRAIN}


class ResidualVQ(nn.Module):
    """`vector_quantize.py:179` eval-mode signature: x [B,C,T] -> (quantized [B,C,T], num_replaces, loss[, indices [B,n,T]])."""

    def __init__(self, deploy: streaming.ResidualVQ):
        super().__init__()
        self._vq = deploy
        self.num_quantizers = len(deploy.layers)

    def forward(self, x: Tensor, n: tp.Optional[int] = None, return_indices: bool = False):
        n = self.num_quantizers if n is None else n
        assert 1 <= n <= self.num_quantizers, \
            f"'n' must be in range of 1 <= n <= {self.num_quantizers}"  # vector_quantize.py:213
        xt = x.transpose(1, 2).contiguous()                       # [B,T,C], vector_quantize.py:206
        idx, q = self._vq.quantize(xt, n)
        quantized_out = q.transpose(1, 2)
        loss = torch.nn.functional.mse_loss(x, quantized_out)      # vector_quantize.py:235
        num_replaces = np.zeros(self.num_quantizers, dtype=np.int64)  # eval: no dead-code replacement
        if return_indices:
            return quantized_out, num_replaces, loss, idx.permute(1, 0, 2)  # [B,n,T]
        return quantized_out, num_replaces, loss


class SEANetEncoder(nn.Module):
    """`modules/seanet.py:249` call shape: x [B,1,T] -> [B,C,ceil(T/hop)] (channel-first)."""

    def __init__(self, deploy: streaming.HILCodec):
        super().__init__()
        object.__setattr__(self, "_deploy", deploy)
        self.hop_length = deploy.cfg.hop

    def forward(self, x: Tensor) -> Tensor:
        d = self._deploy
        if d._core.graph == _lib.HIL_GRAPH_TRAIN:
            return d._core.encode_ragged(x).transpose(1, 2)
        z, _ = d.encoder(x, *d.encoder.initialize_cache(x))
        return z.transpose(1, 2)


class SEANetDecoder(nn.Module):
    """`modules/seanet.py:381` call shape: z [B,C,F] -> [B,1,hop*F]."""

    def __init__(self, deploy: streaming.HILCodec):
        super().__init__()
        object.__setattr__(self, "_deploy", deploy)

    def forward(self, z: Tensor) -> Tensor:
        d = self._deploy
        y, _ = d.decoder(z.transpose(1, 2).contiguous(), *d.decoder.initialize_cache(z))
        return y


class HILCodec(nn.Module):
    """`models/hilcodec/models.py:24` call contract: forward(x[B,1,T], n=None) -> (wav.float(), num_replaces, loss_vq)."""

    def __init__(self, deploy: streaming.HILCodec, graph: str = "deploy"):
        super().__init__()
        if graph not in _GRAPHS:
            raise ValueError(f"Unknown graph: {graph}")
        if graph == "train" and deploy._core.graph != _lib.HIL_GRAPH_TRAIN:
            # same weights, training-graph output scaling; a separate native model so `deploy` stays usable
            w = fold.to_train_graph(deploy._core.weights)
            deploy = streaming.HILCodec.from_weights(w, deploy.cfg.num_quantizers, deploy.sample_rate)
            deploy._core.set_graph(_lib.HIL_GRAPH_TRAIN)
        self.graph = graph
        self.deploy = deploy
        self.encoder = SEANetEncoder(deploy)
        self.quantizer = ResidualVQ(deploy.quantizer)
        self.decoder = SEANetDecoder(deploy)
        self.sample_rate = deploy.sample_rate
        self.channels = deploy.channels

    # -- construction --------------------------------------------------------------------------------------
    @classmethod
    def _from_weights(cls, w, cfg: CodecConfig, graph: str, sample_rate: int) -> "HILCodec":
        d = streaming.HILCodec.from_weights(w, cfg.num_quantizers, sample_rate)
        d._core.set_graph(_GRAPHS[graph])
        return cls(d, graph)

    @classmethod
    def from_pretrained(cls, name: str, graph: str = "deploy") -> "HILCodec":
        """Published `hil_speech` / `hil_music` weights (deployment-folded in the ONNX files)."""
        w = load_pretrained(name)
        return cls._from_weights(fold.to_train_graph(w) if graph == "train" else w, CONFIGS[name], graph, 24_000)

    @classmethod
    def from_training_state_dict(cls, state_dict: tp.Mapping[str, tp.Any], num_quantizers: int, graph: str = "train",
                                 sample_rate: int = 24_000, norm: tp.Optional[str] = None,
                                 norm_kwargs: tp.Optional[tp.Mapping[str, tp.Any]] = None) -> "HILCodec":
        """`models.HILCodec.state_dict()` (weight-norm or weight-standardisation parametrised -- `norm`, `norm_kwargs`
        as given to the reference constructor, `models.py:49-50` -- un-merged scales) -> a servable model."""
        cfg = CodecConfig(num_quantizers=num_quantizers)
        return cls._from_weights(checkpoint.deployment_weights(state_dict, cfg, graph, norm, norm_kwargs), cfg, graph,
                                 sample_rate)

    @classmethod
    def from_checkpoint(cls, path: str, num_quantizers: int, graph: str = "train", sample_rate: int = 24_000,
                        norm: tp.Optional[str] = None, norm_kwargs: tp.Optional[tp.Mapping[str, tp.Any]] = None,
                        allow_pickle: bool = False) -> "HILCodec":
        """A `{epoch:05d}.pth` written by the training wrapper (`wrapper.py:428-444`)."""
        cfg = CodecConfig(num_quantizers=num_quantizers)
        return cls._from_weights(checkpoint.load_checkpoint(path, cfg, graph, norm, norm_kwargs, allow_pickle), cfg, graph,
                                 sample_rate)

    # -- calls -----------------------------------------------------------------------------------------------
    def encode(self, x: Tensor) -> Tensor:
        return self.encoder(x)

    def decode(self, q: Tensor) -> Tensor:
        return self.decoder(q)

    @torch.no_grad()
    def forward(self, x: Tensor, n: tp.Optional[int] = None):
        z = self.encoder(x)
        q, num_replaces, loss_vq = self.quantizer(z, n)
        return self.decoder(q).float(), num_replaces, loss_vq

    def remove_weight_reparameterizations(self) -> None:
        """models.py:120-124.  Weights are folded when loaded, so nothing is left to do."""
        return None
