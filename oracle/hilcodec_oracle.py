"""CPU oracle for the HILCodec encode -> RVQ -> decode forward path.

TEST INFRASTRUCTURE ONLY.  This file is the checker the CUDA path is compared
against; it is never imported by the product package `hilcodec_b200`.  Only
`tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` may import it.

It is a functional restatement (no nn.Module, weights passed as a dict) of the
reference's deployment graph, computed with torch CPU ops so the arithmetic
backend (ATen/oneDNN fp32) is the one the reference itself runs on:

  encoder   models/hilcodec/streaming.py:482-517  (Encoder.forward)
  SpecBlock models/hilcodec/streaming.py:346-365, causal_layers.py:135-144
  ResBlock  models/hilcodec/streaming.py:252-275, DWSBlock :189-192
  causal convs  models/hilcodec/causal_layers.py:160-165 (Conv1d), :183-188 (ConvTranspose1d)
  RVQ       models/hilcodec/streaming.py:51-68 (codebook search), :89-100 (residual loop)
  dequant   models/hilcodec/streaming.py:129-131, :148-157
  decoder   models/hilcodec/streaming.py:619-648  (Decoder.forward)

Weights are the *folded* deployment weights (after
`remove_weight_reparameterizations`, streaming.py:740-747), i.e. exactly the
initializers of the published ONNX graphs; keys are
`encoder.<state_dict key>`, `decoder.<state_dict key>`,
`quantizer.layers.{i}.embed`.

Parity pin (tests/test_oracle.py): with the published `hil_speech` weights the
oracle reproduces all 18 368 golden indices of `onnx/hil_speech_quantized.npy`
and `onnx/hil_speech_output.wav` within one int16 LSB, and it matches the
reference's own `streaming.py` classes (imported in the dev container) on
random weights, one-shot and frame-by-frame.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Params = Dict[str, Tensor]


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
        return tuple(reversed(self.strides))  # streaming.py:388


# --------------------------------------------------------------------------- primitives

def elu(x: Tensor) -> Tensor:
    return F.elu(x, alpha=1.0)


def causal_conv1d(x: Tensor, cache: Tensor, w: Tensor, b: Optional[Tensor],
                  stride: int, groups: int) -> Tuple[Tensor, Tensor]:
    """causal_layers.py:160-165: no implicit padding, history comes from `cache`."""
    pad = cache.shape[2]
    xin = torch.cat((cache, x), dim=2)
    new_cache = xin[:, :, xin.shape[2] - pad:]
    return F.conv1d(xin, w, b, stride=stride, groups=groups), new_cache


def causal_conv_transpose1d(x: Tensor, cache: Tensor, w: Tensor, stride: int) -> Tuple[Tensor, Tensor]:
    """causal_layers.py:168-188 for the depthwise k=2s case: padding=s, output_padding=0."""
    k = w.shape[2]
    causal_padding = (k - 1) // stride
    padding = causal_padding * stride
    output_padding = stride - 1 + padding - (k - 1)
    xin = torch.cat((cache, x), dim=2)
    new_cache = xin[:, :, xin.shape[2] - causal_padding:]
    y = F.conv_transpose1d(xin, w, None, stride=stride, padding=padding,
                           output_padding=output_padding, groups=w.shape[0])
    return y, new_cache


def stft_logmag(wav_window: Tensor, w_dft: Tensor, hop: int) -> Tensor:
    """CausalSTFT (causal_layers.py:135-144) then clamp/log (streaming.py:351)."""
    s = F.conv1d(wav_window, w_dft, None, stride=hop)
    b, c, t = s.shape
    s = s.view(b, 2, c // 2, t)
    mag = s.square().sum(dim=1).sqrt()
    return mag.clamp_min(1e-5).log()


def dws_unit(x: Tensor, cache: Tensor, p: Params, prefix: str) -> Tuple[Tensor, Tensor]:
    """DWSBlock.forward streaming.py:189-192: ELU -> 1x1 (no bias) -> depthwise k5 (+bias)."""
    w_pw = p[prefix + "pointwise.1.weight"]
    w_dw = p[prefix + "depthwise.weight"]
    b_dw = p.get(prefix + "depthwise.bias")
    y = F.conv1d(elu(x), w_pw)
    return causal_conv1d(y, cache, w_dw, b_dw, 1, w_dw.shape[0])


def res_block(x: Tensor, caches: Sequence[Tensor], p: Params, prefix: str,
              pre_scale: float) -> Tuple[Tensor, List[Tensor]]:
    """ResBlock.forward streaming.py:252-275 with merged (folded) residual scale."""
    skip = x
    u = x * pre_scale
    out: List[Tensor] = []
    for i in range(2):
        u, c = dws_unit(u, caches[i], p, f"{prefix}block.{i}.")
        out.append(c)
    return u + skip, out


# --------------------------------------------------------------------------- encoder

def encoder_cache_shapes(cfg: CodecConfig, batch: int) -> List[Tuple[int, int, int]]:
    """Encoder.initialize_cache streaming.py:458-470."""
    n_fft_post = cfg.n_fft_base * 2 ** len(cfg.strides)
    shapes = [(batch, 1, n_fft_post - 1)]
    c = cfg.channels_enc
    for r in cfg.enc_ratios:
        shapes += [(batch, c, cfg.kernel_size - 1)] * (2 * cfg.n_residual_enc)
        shapes.append((batch, 2 * c, 2 * r - 1 - (r - 1)))
        c *= 2
    shapes.append((batch, c, cfg.kernel_size - 1))
    return shapes


def decoder_cache_shapes(cfg: CodecConfig, batch: int) -> List[Tuple[int, int, int]]:
    """Decoder.initialize_cache streaming.py:599-607."""
    c = cfg.channels_dec * 2 ** len(cfg.strides)
    shapes = [(batch, c, cfg.kernel_size - 1)]
    for r in cfg.strides:
        shapes.append((batch, c, (2 * r - 1) // r))
        shapes += [(batch, c // 2, cfg.kernel_size - 1)] * (2 * cfg.n_residual_dec)
        c //= 2
    shapes.append((batch, c, cfg.kernel_size - 1))
    return shapes


def zero_caches(shapes, dtype=torch.float32) -> List[Tensor]:
    return [torch.zeros(s, dtype=dtype) for s in shapes]


def encoder_forward(cfg: CodecConfig, p: Params, x: Tensor,
                    caches: Optional[Sequence[Tensor]] = None) -> Tuple[Tensor, List[Tensor]]:
    """Encoder.forward streaming.py:482-517.  x [B,1,T] -> z [B,T/hop,dim], 22 caches."""
    if caches is None:
        caches = zero_caches(encoder_cache_shapes(cfg, x.shape[0]), x.dtype)
    e = "encoder."
    out: List[Tensor] = []
    wav_cache_len = caches[0].shape[2]
    wav = torch.cat((caches[0], x), dim=2)
    out.append(wav[:, :, wav.shape[2] - wav_cache_len:])
    kpre = cfg.kernel_size
    h = F.conv1d(wav[:, :, wav_cache_len - (kpre - 1):], p[e + "conv_pre.weight"], p[e + "conv_pre.bias"])
    idx = 1
    stride = 1
    post_scale = (1 + cfg.n_residual_enc * cfg.res_scale_enc ** 2) ** -0.5
    for s, r in enumerate(cfg.enc_ratios):
        n_fft = cfg.n_fft_base * 2 ** s
        y = stft_logmag(wav[:, :, wav_cache_len - (n_fft - 1):], p[f"{e}spec_blocks.{s}.spec.weight"], stride)
        h = F.conv1d(y, p[f"{e}spec_blocks.{s}.layer.weight"], p[f"{e}spec_blocks.{s}.layer.bias"]) + h
        for j in range(cfg.n_residual_enc):
            pre = (1 + (j + 1) * cfg.res_scale_enc ** 2) ** -0.5  # streaming.py:210, idx=j+1 (:418-426)
            h, cs = res_block(h, caches[idx:idx + 2], p, f"{e}blocks.{s}.{j}.", pre)
            out.extend(cs)
            idx += 2
        h = h * post_scale
        h = F.conv1d(elu(h), p[f"{e}downsample_pointwise.{s}.1.weight"])
        h, c = causal_conv1d(h, caches[idx], p[f"{e}downsample_depthwise.{s}.weight"],
                             p[f"{e}downsample_depthwise.{s}.bias"], r, h.shape[1])
        out.append(c)
        idx += 1
        stride *= r
    y = stft_logmag(wav, p[e + "spec_post.spec.weight"], stride)
    h = F.conv1d(y, p[e + "spec_post.layer.weight"], p[e + "spec_post.layer.bias"]) + h
    h, c = causal_conv1d(elu(h), caches[idx], p[e + "conv_post_depthwise.weight"], None, 1, h.shape[1])
    out.append(c)
    h = F.conv1d(h, p[e + "conv_post_pointwise.weight"], p[e + "conv_post_pointwise.bias"])
    h = F.normalize(h, p=2.0, dim=1, eps=1e-12) * (cfg.dim ** 0.5)  # L2Norm streaming.py:284-285
    return h.transpose(1, 2), out


# --------------------------------------------------------------------------- RVQ

def codebook_search(x: Tensor, embed: Tensor) -> Tuple[Tensor, Tensor]:
    """EuclideanCodebook.forward streaming.py:51-68 (argmax of the negated full distance)."""
    b, t, c = x.shape
    flat = x.reshape(b * t, c)
    et = embed.t()
    dist = -(flat.pow(2).sum(1, keepdim=True) - 2 * flat @ et + et.pow(2).sum(0, keepdim=True))
    ind = dist.max(dim=-1).indices.view(b, t)
    return F.embedding(ind, embed), ind


def rvq_encode(cfg: CodecConfig, p: Params, z: Tensor, n: int) -> Tensor:
    """ResidualVQ.forward streaming.py:89-100.  z [B,F,dim] -> indices [n,B,F] int64."""
    assert 1 <= n <= cfg.num_quantizers
    residual = z
    indices = []
    for i in range(n):
        q, ind = codebook_search(residual, p[f"quantizer.layers.{i}.embed"])
        residual = residual - q
        indices.append(ind)
    return torch.stack(indices, dim=0)


def rvq_decode(cfg: CodecConfig, p: Params, indices: Tensor, n: int) -> Tensor:
    """Dequantizer.forward streaming.py:148-157: in-order fp32 sum of gathered rows."""
    out = torch.zeros(1, dtype=p["quantizer.layers.0.embed"].dtype)
    for i in range(n):
        out = out + F.embedding(indices[i], p[f"quantizer.layers.{i}.embed"])
    return out


def rvq_margins(cfg: CodecConfig, p: Params, z: Tensor, n: int) -> Tuple[Tensor, Tensor]:
    """float64 re-run of the search along the *given* float32 path: returns the
    indices and the relative gap (d2-d1)/d1 between best and runner-up distance
    of every decision, for the near-tie policy of the parity tests."""
    residual = z.double()
    inds, gaps = [], []
    for i in range(n):
        e = p[f"quantizer.layers.{i}.embed"].double()
        flat = residual.reshape(-1, residual.shape[-1])
        d = (flat.pow(2).sum(1, keepdim=True) - 2 * flat @ e.t() + e.pow(2).sum(1)[None])
        top = torch.topk(d, 2, dim=1, largest=False)
        ind = top.indices[:, 0].view(z.shape[0], z.shape[1])
        gaps.append(((top.values[:, 1] - top.values[:, 0]) / top.values[:, 0].clamp_min(1e-30)).view_as(ind))
        inds.append(ind)
        residual = residual - F.embedding(ind, e)
    return torch.stack(inds), torch.stack(gaps)


# --------------------------------------------------------------------------- decoder

def decoder_forward(cfg: CodecConfig, p: Params, q: Tensor,
                    caches: Optional[Sequence[Tensor]] = None, train_graph: bool = False) -> Tuple[Tensor, List[Tensor]]:
    """Decoder.forward streaming.py:619-648.  q [B,F,dim] -> wav [B,1,hop*F], 30 caches.

    Deploy-path quirks kept on purpose: ResBlock pre_scale is 1.0 (streaming.py:576-583
    never passes idx) and conv_post.bias is not scaled by wav_std (:609-617).

    `train_graph=True` restates SEANetDecoder (modules/seanet.py:381-479) instead: ResBlock j of every
    stage uses pre_scale (1 + j * res_scale^2)^-0.5 (seanet.py:443-451 passes idx=j), and `p` must be
    folded for the training graph (conv_post.bias * wav_std, hilcodec_b200/fold.py graph="train")."""
    if caches is None:
        caches = zero_caches(decoder_cache_shapes(cfg, q.shape[0]), q.dtype)
    d = "decoder."
    out: List[Tensor] = []
    h = F.conv1d(q.transpose(1, 2), p[d + "conv_pre_pointwise.weight"])
    h, c = causal_conv1d(h, caches[0], p[d + "conv_pre_depthwise.weight"],
                         p[d + "conv_pre_depthwise.bias"], 1, h.shape[1])
    out.append(c)
    idx = 1
    post_scale = (1 + cfg.n_residual_dec * cfg.res_scale_dec ** 2) ** -0.5
    for i, r in enumerate(cfg.strides):
        h, c = causal_conv_transpose1d(elu(h), caches[idx], p[f"{d}upsample_depthwise.{i}.weight"], r)
        h = F.conv1d(h, p[f"{d}upsample_pointwise.{i}.weight"], p[f"{d}upsample_pointwise.{i}.bias"])
        out.append(c)
        idx += 1
        for j in range(cfg.n_residual_dec):
            pre = (1 + j * cfg.res_scale_dec ** 2) ** -0.5 if train_graph else 1.0
            h, cs = res_block(h, caches[idx:idx + 2], p, f"{d}blocks.{i}.{j}.", pre)
            out.extend(cs)
            idx += 2
        h = h * post_scale
    h, c = causal_conv1d(elu(h), caches[idx], p[d + "conv_post.weight"], p[d + "conv_post.bias"], 1, 1)
    out.append(c)
    return torch.tanh(h), out


# --------------------------------------------------------------------------- whole path

def codec_forward(cfg: CodecConfig, p: Params, x: Tensor, n: int,
                  enc_caches=None, dec_caches=None):
    """encode -> RVQ -> dequant -> decode, the four-call flow of
    scripts/HILCodec Onnx.ipynb cell 3 / test_onnx.py:75-135."""
    z, ce = encoder_forward(cfg, p, x, enc_caches)
    idx = rvq_encode(cfg, p, z, n)
    q = rvq_decode(cfg, p, idx, n)
    y, cd = decoder_forward(cfg, p, q, dec_caches)
    return {"z": z, "indices": idx, "q": q, "wav": y, "enc_caches": ce, "dec_caches": cd}


# --------------------------------------------------------------------------- training graph
# models/hilcodec/models.py:111-118 = SEANetEncoder -> ResidualVQ(channel_last=False) -> SEANetDecoder, eval mode,
# restated on FOLDED weights (training-graph fold).  Differences from the deployment graph above:
#   * every causal conv pads itself: left (k-1) - (s-1) zeros, right "extra padding" so that the last window is full
#     (modules/conv.py:61-68, :222-236) => any T >= 1 is accepted and T_out = ceil(T_in / stride) per layer;
#   * the STFT pads n_fft-1 zeros on the left (modules/conv.py:348-358) and clamps the power at 1e-12 before the sqrt
#     (invisible behind the log's clamp at 1e-5);
#   * codebook search drops the sum(x^2) term and takes argmin (vector_quantize.py:132-153);
#   * decoder pre_scale / conv_post bias as described in decoder_forward(train_graph=True).

def _extra_padding(length: int, kernel: int, stride: int, padding_total: int) -> int:
    """get_extra_padding_for_conv1d, modules/conv.py:61-68."""
    n_frames = (length - kernel + padding_total) / stride + 1
    ideal = (math.ceil(n_frames) - 1) * stride + (kernel - padding_total)
    return ideal - length


def sconv1d_causal(x: Tensor, w: Tensor, b: Optional[Tensor], stride: int, groups: int) -> Tensor:
    """SConv1d.forward modules/conv.py:222-236 (causal, pad_mode constant, dilation 1)."""
    k = w.shape[2]
    padding_total = (k - 1) - (stride - 1)
    extra = _extra_padding(x.shape[2], k, stride, padding_total)
    return F.conv1d(F.pad(x, (padding_total, extra)), w, b, stride=stride, groups=groups)


def stft_logmag_train(wav: Tensor, w_dft: Tensor, hop: int) -> Tensor:
    """CausalSTFT.forward modules/conv.py:348-358 + SpecBlock compression seanet.py:231-232."""
    s = F.conv1d(F.pad(wav, (w_dft.shape[2] - 1, 0)), w_dft, None, stride=hop)
    b, c, t = s.shape
    mag = s.view(b, 2, c // 2, t).square().sum(dim=1).clamp_min(1e-12).sqrt()
    return mag.clamp_min(1e-5).log()


def encoder_forward_train(cfg: CodecConfig, p: Params, x: Tensor) -> Tensor:
    """SEANetEncoder.forward seanet.py:368-378.  x [B,1,T] (any T >= 1) -> z [B,dim,ceil(T/hop)] (channel-first)."""
    e = "encoder."
    wav = x
    h = sconv1d_causal(x, p[e + "conv_pre.weight"], p[e + "conv_pre.bias"], 1, 1)
    stride = 1
    post_scale = (1 + cfg.n_residual_enc * cfg.res_scale_enc ** 2) ** -0.5

    def dws(u: Tensor, prefix: str) -> Tensor:
        w_dw = p[prefix + "depthwise.weight"]
        u = F.conv1d(elu(u), p[prefix + "pointwise.1.weight"])
        return sconv1d_causal(u, w_dw, p.get(prefix + "depthwise.bias"), 1, w_dw.shape[0])

    for s, r in enumerate(cfg.enc_ratios):
        y = stft_logmag_train(wav, p[f"{e}spec_blocks.{s}.spec.weight"], stride)
        h = h + F.conv1d(y, p[f"{e}spec_blocks.{s}.layer.weight"], p[f"{e}spec_blocks.{s}.layer.bias"])
        for j in range(cfg.n_residual_enc):
            pre = (1 + (j + 1) * cfg.res_scale_enc ** 2) ** -0.5  # seanet.py:295-296, idx=j (1-based) with a spec branch
            u = h * pre
            for i in range(2):
                u = dws(u, f"{e}blocks.{s}.{j}.block.{i}.")
            h = u + h
        h = F.conv1d(elu(h * post_scale), p[f"{e}downsample_pointwise.{s}.1.weight"])
        h = sconv1d_causal(h, p[f"{e}downsample_depthwise.{s}.weight"], p[f"{e}downsample_depthwise.{s}.bias"],
                           r, h.shape[1])
        stride *= r
    y = stft_logmag_train(wav, p[e + "spec_post.spec.weight"], stride)
    h = h + F.conv1d(y, p[e + "spec_post.layer.weight"], p[e + "spec_post.layer.bias"])
    w_dw = p[e + "conv_post_depthwise.weight"]
    h = sconv1d_causal(elu(h), w_dw, None, 1, w_dw.shape[0])
    h = F.conv1d(h, p[e + "conv_post_pointwise.weight"], p[e + "conv_post_pointwise.bias"])
    return F.normalize(h, p=2.0, dim=1, eps=1e-12) * (cfg.dim ** 0.5)  # L2Norm seanet.py:152-163


def codebook_search_train(x: Tensor, embed: Tensor) -> Tuple[Tensor, Tensor]:
    """EuclideanCodebook.forward (eval) models/hilcodec/vector_quantize.py:132-153: no sum(x^2), argmin."""
    b, t, c = x.shape
    flat = x.reshape(b * t, c)
    et = embed.t()
    distance = - 2 * flat @ et + et.pow(2).sum(0, keepdim=True)
    ind = distance.min(dim=-1).indices.view(b, t)
    return F.embedding(ind, embed), ind


def rvq_forward_train(cfg: CodecConfig, p: Params, x: Tensor, n: Optional[int] = None):
    """ResidualVQ.forward (eval) vector_quantize.py:199-243.  x [B,dim,F] -> (quantized [B,dim,F], loss, indices [B,n,F])."""
    high = cfg.num_quantizers if n is None else n
    assert 1 <= high <= cfg.num_quantizers
    residual = x.transpose(1, 2)
    indices = []
    out = None
    for i in range(high):
        q, ind = codebook_search_train(residual, p[f"quantizer.layers.{i}.embed"])
        indices.append(ind)
        residual = residual - q
        out = q if out is None else out + q
    out = out.transpose(1, 2)
    return out, F.mse_loss(x, out), torch.stack(indices, dim=1)


def codec_forward_train(cfg: CodecConfig, p: Params, x: Tensor, n: Optional[int] = None):
    """HILCodec.forward models/hilcodec/models.py:111-118 (eval).  `p`: weights folded with graph="train"."""
    z = encoder_forward_train(cfg, p, x)
    q, loss, idx = rvq_forward_train(cfg, p, z, n)
    y, _ = decoder_forward(cfg, p, q.transpose(1, 2), None, train_graph=True)
    return {"z": z, "q": q, "indices": idx, "loss_vq": loss, "wav": y}


# --------------------------------------------------------------------------- generic RVQ (modules/vector_quantize.py)

def generic_rvq_forward(embeds: Sequence[Tensor], x: Tensor, n: Optional[int] = None, channel_last: bool = False):
    """`ResidualVQ.forward` (eval) of the GENERIC quantizer, modules/vector_quantize.py:490-516, with
    `VectorQuantize.forward` :400-419 and `EuclideanCodebook.forward` :141-160 inlined: per stage the residual is
    (transposed to [B,T,C] unless channel_last and) searched with the same negated full distance as the deployment
    codebook, `quantized_out = 0. + q_0 + q_1 + ...` in the INPUT layout.  Returns (quantized_out, loss, indices
    [n,B,T]) -- the reference returns (quantized_out, num_replaces = zeros, loss); the indices are extra."""
    high = len(embeds) if n is None else n
    assert 1 <= high <= len(embeds)
    quantized_out = 0.
    residual = x
    indices = []
    for embed in embeds[:high]:
        r = residual if channel_last else residual.transpose(1, 2)
        q, ind = codebook_search(r, embed)          # :151-160, same expression as streaming.py:58-66
        if not channel_last:
            q = q.transpose(1, 2)
        indices.append(ind)
        residual = residual - q
        quantized_out = quantized_out + q
    return quantized_out, F.mse_loss(x, quantized_out), torch.stack(indices)


def to_dtype(p: Params, dtype) -> Params:
    return {k: v.to(dtype) for k, v in p.items()}
