"""Load-time weight preparation: fold a *training-format* streaming state dict into the
deployment tensors the kernels consume.

Restates `HILCodec.remove_weight_reparameterizations` (streaming.py:740-747):
  * weight_norm removal, w = g * v / ||v||  (norm over all dims but 0; both the legacy
    `weight_g`/`weight_v` and the `parametrizations.weight.original{0,1}` spellings),
  * `Encoder.merge_scaling`  (streaming.py:472-480): conv_pre.weight /= wav_std,
  * `SpecBlock.merge_scaling` (streaming.py:321-344): W/std, bias = -(sum W) * mean/std,
    both times res_scale * scale_param,
  * `ResBlock.merge_scaling`  (streaming.py:240-250): last depthwise weight/bias times
    res_scale * res_scale_param,
  * `Decoder.merge_scaling`  (streaming.py:609-617): conv_post.weight *= wav_std (bias untouched).
  * weight standardisation removal (`modules/weight_standardization.py:30-41` compute_weight, `:137-146` remove;
    selected in the training graph with `norm="weight_standardization"`, `modules/conv.py:36-37`):
    w = (g * scale) * (v - mean(v)) * rsqrt(clamp(var(v) * fan_in, eps)), statistics over every axis but `dim`
    (biased variance), `g` absent when `learnable_gain=False`, `scale` absent when not given.  The state-dict keys
    are `weight_v` / `weight_g` / `weight_scale` -- the first two collide with legacy weight_norm, so the
    normalisation is named by the caller (`norm=`), and auto-detected only from `weight_scale` or a `weight_v`
    without `weight_g`.
A dict that is already folded (no reparametrisation keys) passes through unchanged.

`graph="train"` folds for the TRAINING graph instead (models/hilcodec/modules/seanet.py): identical except
that the decoder ends with `Scale(wav_std)` AFTER conv + bias (seanet.py:464-466), so `conv_post.bias` is
scaled by wav_std as well.  `to_train_graph()` applies that one difference to weights
that are already folded for deployment.
"""
from __future__ import annotations

import re
import typing as tp
from collections import OrderedDict

import numpy as np
import torch
from torch import Tensor

from .weights import WAV_STD, CodecConfig

SPEC_MEANS = [-4.554, -4.315, -4.021, -3.726, -3.477]  # streaming.py:383
SPEC_STDS = [2.830, 2.837, 2.817, 2.796, 2.871]        # streaming.py:384


NORMS = ("weight_norm", "weight_standardization")


def _needs_fold(sd: tp.Mapping[str, Tensor]) -> bool:
    return any(k.endswith(("weight_g", "weight_v", "weight_scale", "res_scale_param", "scale_param",
                           "parametrizations.weight.original0")) for k in sd)


def detect_norm(sd: tp.Mapping[str, tp.Any]) -> tp.Optional[str]:
    """"weight_standardization" when the keys can only come from it (`weight_scale`, or a `weight_v` whose gain is
    not learnable), "weight_norm" for the parametrize spelling, None when the keys do not tell (`weight_g` +
    `weight_v` are written by both)."""
    keys = set(sd)
    if any(k.endswith("weight_scale") for k in keys):
        return "weight_standardization"
    if any(k.endswith("weight_v") and k[:-1] + "g" not in keys for k in keys):
        return "weight_standardization"
    if any(k.endswith("parametrizations.weight.original0") for k in keys):
        return "weight_norm"
    return None


def standardize_weight(v: Tensor, g: tp.Optional[Tensor] = None, scale: tp.Optional[Tensor] = None,
                       dim: tp.Union[int, tp.Sequence[int]] = 0, eps: float = 1e-7) -> Tensor:
    """`WeightStandardization.compute_weight` (modules/weight_standardization.py:30-41), same operation order."""
    v = v.float()
    if isinstance(dim, int):
        dim = (dim + v.dim() if dim < -1 else dim,)   # weight_standardization.py:65-68 (sic: -1 is left as is)
    axes = [a for a in range(v.dim()) if a not in tuple(dim)]
    fan_in = 1.0
    for a in axes:
        fan_in *= v.size(a)
    var, mean = torch.var_mean(v, dim=axes, unbiased=False, keepdim=True)
    w = (v - mean) * torch.rsqrt(torch.clamp(var * fan_in, min=eps))
    if g is not None:
        g = g.float()
        if scale is not None:
            g = g * scale.float()
        w = g * w
    return w


def _remove_weight_standardization(sd: "OrderedDict[str, Tensor]", dim: tp.Union[int, tp.Sequence[int]] = 0,
                                   eps: float = 1e-7, **_unused) -> "OrderedDict[str, Tensor]":
    """`WeightStandardization.remove` (modules/weight_standardization.py:137-146) on every `weight_v` of a dict."""
    out: "OrderedDict[str, Tensor]" = OrderedDict()
    for k, v in sd.items():
        if k.endswith("weight_v"):
            base = k[:-len("weight_v")]
            out[base + "weight"] = standardize_weight(v, sd.get(base + "weight_g"), sd.get(base + "weight_scale"), dim, eps)
        elif k.endswith(("weight_g", "weight_scale")) and k[:k.rindex("weight_")] + "weight_v" in sd:
            continue
        else:
            out[k] = v
    return out


def _remove_weight_norm(sd: "OrderedDict[str, Tensor]") -> "OrderedDict[str, Tensor]":
    out: "OrderedDict[str, Tensor]" = OrderedDict()
    for k, v in sd.items():
        if k.endswith("weight_g") or k.endswith("parametrizations.weight.original0"):
            base = k[:-len("weight_g")] if k.endswith("weight_g") else k[:-len("parametrizations.weight.original0")]
            vk = base + ("weight_v" if k.endswith("weight_g") else "parametrizations.weight.original1")
            g, vv = v.float(), sd[vk].float()
            norm = vv.reshape(vv.shape[0], -1).norm(dim=1).reshape(-1, *([1] * (vv.dim() - 1)))
            out[base + "weight"] = vv * (g / norm)  # torch._weight_norm(v, g, dim=0)
        elif k.endswith("weight_v") or k.endswith("parametrizations.weight.original1"):
            continue
        else:
            out[k] = v
    return out


def to_train_graph(weights: tp.Mapping[str, tp.Any], part: str = "") -> "OrderedDict[str, tp.Any]":
    """Deployment-folded weights -> training-graph weights: conv_post.bias *= wav_std (seanet.py:464-466)."""
    out = OrderedDict(weights)
    key = ("decoder." if part == "" else "") + "conv_post.bias"
    if key in out:
        out[key] = out[key] * float(WAV_STD) if isinstance(out[key], Tensor) else (out[key] * np.float32(WAV_STD))
    return out


def fold_state_dict(sd: "OrderedDict[str, Tensor]", cfg: CodecConfig, part: str = "",
                    graph: str = "deploy", norm: tp.Optional[str] = None,
                    norm_kwargs: tp.Optional[tp.Mapping[str, tp.Any]] = None) -> "OrderedDict[str, Tensor]":
    """`part` is "" for a HILCodec-level dict (keys start with encoder./decoder./...),
    or "encoder" / "decoder" / "quantizer" for a sub-module dict.  `norm` / `norm_kwargs` are the training model's
    constructor arguments of the same names (`models/hilcodec/models.py:49-50`); `norm=None` means: what the keys
    say (`detect_norm`), else weight_norm (the default of the reference and of both published configs)."""
    if graph not in ("deploy", "train"):
        raise ValueError(f"Unknown graph: {graph}")
    sd = OrderedDict((k, torch.as_tensor(v).detach().cpu()) for k, v in sd.items())
    if not _needs_fold(sd):
        # already folded for deployment: the training graph differs only in conv_post.bias (quirk 2)
        return to_train_graph(sd, part) if graph == "train" and part in ("", "decoder") else sd
    if norm is None:
        norm = detect_norm(sd) or "weight_norm"
    if norm not in NORMS:
        raise ValueError(f"Unknown norm: {norm}")  # causal_layers.py:203
    if norm == "weight_standardization":
        sd = _remove_weight_standardization(sd, **dict(norm_kwargs or {}))
    else:
        sd = _remove_weight_norm(sd)
    enc = "encoder." if part == "" else ("" if part == "encoder" else None)
    dec = "decoder." if part == "" else ("" if part == "decoder" else None)
    n_stage = len(cfg.strides)

    def scale_conv(prefix: str, s: Tensor) -> None:
        sd[prefix + "weight"] = sd[prefix + "weight"].float() * s
        if prefix + "bias" in sd:
            sd[prefix + "bias"] = sd[prefix + "bias"].float() * s

    for p, rs in ((enc, cfg.res_scale_enc), (dec, cfg.res_scale_dec)):
        if p is None:
            continue
        for k in [k for k in sd if k.startswith(p) and k.endswith("res_scale_param")]:
            block = k[:-len("res_scale_param")]
            scale_conv(block + "block.1.depthwise.", rs * sd.pop(k).float())
    if enc is not None:
        if enc + "conv_pre.weight" in sd:
            sd[enc + "conv_pre.weight"] = sd[enc + "conv_pre.weight"].float() / WAV_STD
        specs = [(f"{enc}spec_blocks.{i}.", i) for i in range(n_stage)] + [(enc + "spec_post.", n_stage)]
        for sp, i in specs:
            if sp + "scale_param" not in sd:
                continue
            mean, std = SPEC_MEANS[min(i, 4)], SPEC_STDS[min(i, 4)]
            w = sd[sp + "layer.weight"].float()
            bias2 = w.sum((1, 2)) * (-mean / std)
            w = w / std
            b = sd[sp + "layer.bias"].float() + bias2 if sp + "layer.bias" in sd else bias2
            scale = cfg.res_scale_enc * sd.pop(sp + "scale_param").float()
            sd[sp + "layer.weight"] = w * scale
            sd[sp + "layer.bias"] = b * scale
    if dec is not None and dec + "conv_post.weight" in sd:
        sd[dec + "conv_post.weight"] = sd[dec + "conv_post.weight"].float() * WAV_STD
        if graph == "train" and dec + "conv_post.bias" in sd:
            sd[dec + "conv_post.bias"] = sd[dec + "conv_post.bias"].float() * WAV_STD
    return sd
