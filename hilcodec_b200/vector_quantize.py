"""Drop-in for the reference's GENERIC residual vector quantizer, `modules/vector_quantize.py` -- the class
`BASELINE.json:north_star` names ("modules.vector_quantize.ResidualVectorQuantizer"; the class is spelled
`ResidualVQ`, `modules/vector_quantize.py:471`) and the one the EnCodec baseline model constructs
(`models/encodec/models.py:16`).  Same constructor keywords, buffers / `state_dict()` keys and eval-mode
`forward()` contracts:

  EuclideanCodebook(dim, codebook_size, ...)   :76    forward(x[..., dim]) -> (quantize[..., dim], num_replace)      :141-195
  VectorQuantize(commitment, use_shape_gain, channel_last, gradient_flow, **codebook_kwargs)  :376
        forward(x, calculate_commitment_loss=False) -> (quantize, num_replace, commit_loss | None)                  :400-419
  ResidualVQ(num_quantizers, dropout, dropout_index, **vector_quantize_kwargs)  :471
        forward(x[B,C,T] | x[B,T,C] if channel_last, n=None) -> (quantized_out, num_replaces, loss)                 :490-516

The search itself -- `argmax(-(|x|^2 - 2 x.E^T + |E|^2))`, first index on ties (:151-157), `residual -= E[idx]`,
`quantized_out = ((0 + E_0[i_0]) + E_1[i_1]) + ...` in stage order -- runs in the library's one-kernel RVQ
(`hil_rvq_encode`, csrc/rvq.cu), the same kernel the deployment `streaming.ResidualVQ` drop-in uses; the formula
of this class is the deployment one (SURVEY.md section 8a rows Q'' / V'').  Out of scope (training only, SURVEY
section 2 row 6): the EMA codebook update, k-means initialisation, dead-code replacement, the straight-through
term and the `random` layer dropout -- all of them live behind `self.training`, which raises here -- and the
shape-gain codebooks (`use_shape_gain=True`: unused by HILCodec, `models/hilcodec/models.py:101-106`).

Tensors must be CUDA float32 (no CPU path); `dim` must be 128 (the only width the kernel is built for).
"""
from __future__ import annotations

import typing as tp

import numpy as np
import torch
import torch.nn.functional as F
from torch import Tensor, nn

from .streaming import _NativeCodec, _require_cuda
from .weights import CodecConfig

_TRAINING_MSG = ("hilcodec_b200.vector_quantize runs the eval-mode forward only (call .eval()): the EMA codebook "
                 "update / k-means init / code expiry of training mode are out of scope")


class _Search:
    """Native RVQ over a list of `embed` buffers; re-uploads the codebooks when a buffer is replaced or written."""

    def __init__(self, dim: int, codebook_size: int, num_quantizers: int):
        if dim != 128:
            raise NotImplementedError("hilcodec_b200: the RVQ kernel is built for dim=128 (both published configs)")
        self.core = _NativeCodec(CodecConfig(dim=dim, codebook_size=codebook_size, num_quantizers=num_quantizers))
        self._seen: tp.Optional[tp.Tuple[tp.Tuple[int, int], ...]] = None

    def sync(self, embeds: tp.Sequence[Tensor]) -> None:
        stamp = tuple((e.data_ptr(), e._version) for e in embeds)
        if stamp != self._seen:
            self.core.set_weights({f"quantizer.layers.{i}.embed": e for i, e in enumerate(embeds)})
            self._seen = stamp

    def quantize(self, x_btc: Tensor, embeds: tp.Sequence[Tensor], n: int) -> tp.Tuple[Tensor, Tensor]:
        self.sync(embeds)
        return self.core.rvq_encode(x_btc, n, with_sum=True)


class EuclideanCodebook(nn.Module):
    """`modules/vector_quantize.py:76`.  Buffers as in the reference (:99-102): initted, embed, ema_embed, ema_num."""

    def __init__(self, dim: int, codebook_size: int, kmeans_init: bool = False, kmeans_iters: int = 20,
                 decay: float = 0.8, eps: float = 1e-7, ema_num_threshold: float = 0.0, ema_num_initial: float = 1.0):
        super().__init__()
        self.decay = decay
        init_fn = torch.randn if not kmeans_init else torch.zeros
        embed = init_fn(codebook_size, dim)
        self.codebook_size = codebook_size
        self.kmeans_iters = kmeans_iters
        self.eps = eps
        self.ema_num_threshold = ema_num_threshold
        self.ema_num_initial = ema_num_initial
        self.register_buffer("initted", Tensor([not kmeans_init]))
        self.register_buffer("embed", embed)
        self.register_buffer("ema_embed", embed.clone() * ema_num_initial)
        self.register_buffer("ema_num", torch.ones(codebook_size) * ema_num_initial)
        self.distributed = False
        object.__setattr__(self, "_search", _Search(dim, codebook_size, 1))

    @torch.no_grad()
    def forward(self, x: Tensor) -> tp.Tuple[Tensor, int]:
        if self.training:
            raise NotImplementedError(_TRAINING_MSG)
        if not bool(self.initted):
            raise NotImplementedError("hilcodec_b200: codebook not initialised (kmeans_init is a training-time step)")
        _require_cuda(x, "x")
        shape = x.shape
        flat = x.reshape(1, -1, shape[-1])                     # rearrange '... d -> (...) d' (:144)
        _, q = self._search.quantize(flat, [self.embed], 1)
        return q.view(shape), 0


class VectorQuantize(nn.Module):
    """`modules/vector_quantize.py:376`."""

    def __init__(self, commitment: float = 1., use_shape_gain: bool = False, channel_last: bool = False,
                 gradient_flow: bool = True, **kwargs):
        super().__init__()
        if use_shape_gain:
            raise NotImplementedError("hilcodec_b200: ShapeGainCodebook is out of scope (unused by HILCodec)")
        self.commitment = commitment
        self.use_shape_gain = use_shape_gain
        self.channel_last = channel_last
        self._codebook = EuclideanCodebook(**kwargs)
        self.gradient_flow = gradient_flow

    def forward(self, x: Tensor, calculate_commitment_loss: bool = False):
        if self.training:
            raise NotImplementedError(_TRAINING_MSG)
        if not self.channel_last:
            x = x.transpose(1, 2)                              # 'b c t -> b t c' (:401-403)
        quantize, num_replace = self._codebook(x.contiguous())
        commit_loss = F.mse_loss(quantize.detach(), x) * self.commitment if calculate_commitment_loss else None
        if not self.channel_last:
            quantize = quantize.transpose(1, 2)                # 'b t c -> b c t' (:415-417)
        return quantize, num_replace, commit_loss


class ResidualVQ(nn.Module):
    """`modules/vector_quantize.py:471`: Algorithm 1 of arXiv 2107.03312, all stages in ONE kernel launch."""

    def __init__(self, num_quantizers: int, dropout: bool = False, dropout_index: tp.Optional[tp.List[int]] = None,
                 **kwargs):
        super().__init__()
        self.layers = nn.ModuleList([VectorQuantize(gradient_flow=False, **kwargs) for _ in range(num_quantizers)])
        self.dropout = dropout
        if dropout_index is None:
            dropout_index = list(range(1, num_quantizers + 1))
        self.dropout_index = dropout_index
        self.use_shape_gain = self.layers[0].use_shape_gain
        cb = self.layers[0]._codebook
        object.__setattr__(self, "_search", _Search(cb.embed.shape[1], cb.codebook_size, num_quantizers))

    def forward(self, x: Tensor, n: tp.Optional[int] = None) -> tp.Tuple[Tensor, np.ndarray, Tensor]:
        if self.training:
            raise NotImplementedError(_TRAINING_MSG)
        num_replaces = np.zeros(len(self.layers), dtype=np.int64)              # :493 (eval: nothing is replaced)
        if n is not None:
            assert 1 <= n <= len(self.layers), \
                f"'n' must be in range of 1 <= n <= {len(self.layers)}"       # :496-497
            high = n
        else:
            high = len(self.layers)
        _require_cuda(x, "x")
        channel_last = self.layers[0].channel_last
        with torch.no_grad():
            xt = (x if channel_last else x.transpose(1, 2)).contiguous()       # [B,T,C]
            embeds = [layer._codebook.embed for layer in self.layers]
            if not all(bool(layer._codebook.initted) for layer in self.layers[:high]):
                raise NotImplementedError("hilcodec_b200: codebook not initialised (kmeans_init is a training-time step)")
            _, q = self._search.quantize(xt, embeds, high)
            quantized_out = q if channel_last else q.transpose(1, 2)
            loss = F.mse_loss(x, quantized_out)                                 # :512
        return quantized_out, num_replaces, loss
