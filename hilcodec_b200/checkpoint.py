"""Serve a *trained* HILCodec checkpoint: training-graph `state_dict` -> deployment tensors
(SURVEY.md section 8f.2).

The reference does this in a notebook (`scripts/HILCodec Onnx.ipynb` cell 1): it builds
`streaming.HILCodec`, copies every conv of the training model (`models/hilcodec/models.py:24`,
`modules/seanet.py`) into the streaming module that plays the same role, copies the codebooks, and calls
`remove_weight_reparameterizations()` (streaming.py:740-747).  Here the copy is a pure key rename
(`training_to_streaming`) followed by `fold.fold_state_dict`:

  training key (models.HILCodec.state_dict())                   streaming key
  encoder.conv_pre.1.conv.conv.*                                encoder.conv_pre.*
  encoder.blocks.{s}.{j}.block.{1,2,4,5}.conv.conv.*            encoder.blocks.{s}.{j}.block.{0,0,1,1}.{pointwise.1,depthwise}.*
  encoder.blocks.{s}.{j}.res_scale_param                        (same)
  encoder.spec_blocks.{s}.layer.conv.conv.* / .scale_param      encoder.spec_blocks.{s}.layer.* / .scale_param
  encoder.downsample.{s}.{2,3}.conv.conv.*                      encoder.downsample_{pointwise.{s}.1,depthwise.{s}}.*
  encoder.spec_post.layer.conv.conv.* / .scale_param            encoder.spec_post.layer.* / .scale_param
  encoder.conv_post.{1,2}.conv.conv.*                           encoder.conv_post_{depthwise,pointwise}.*
  decoder.model.{0,1}.conv.conv.*                               decoder.conv_pre_{pointwise,depthwise}.*
  decoder.model.{i}.convtr.convtr.* , next .conv.conv.*         decoder.upsample_{depthwise,pointwise}.{stage}.*
  decoder.model.{i}.block.{1,2,4,5}.conv.conv.* (+res_scale_param)  decoder.blocks.{stage}.{j}.block....
  decoder.model.{last conv}.conv.conv.*                         decoder.conv_post.*
  quantizer.layers.{i}.embed                                    quantizer.layers.{i}.embed

(`weight_g` / `weight_v` or `parametrizations.weight.original{0,1}` suffixes travel unchanged; the fixed DFT
bases `spec.weight`, the EMA statistics and `_extra_state` are dropped -- the kernels rebuild the bases.)

`load_checkpoint()` reads the `{epoch:05d}.pth` files written by the training wrapper
(`models/hilcodec/wrapper.py:428-444`: a dict whose "model" entry is that state dict).
"""
from __future__ import annotations

import math
import re
import typing as tp
from collections import OrderedDict

import numpy as np
import torch
from torch import Tensor

from . import fold
from .weights import WAV_STD, CodecConfig, dft_basis, random_weights, tensor_shapes

_PARAM = r"(bias|weight|weight_g|weight_v|weight_scale|parametrizations\.weight\.original[01])"
_BLOCK_SLOT = {1: "block.0.pointwise.1", 2: "block.0.depthwise", 4: "block.1.pointwise.1", 5: "block.1.depthwise"}


def is_training_state_dict(sd: tp.Mapping[str, tp.Any]) -> bool:
    """True for `models.HILCodec.state_dict()` naming (as opposed to `streaming.HILCodec`'s)."""
    return any(k.startswith("decoder.model.") or ".conv.conv." in k for k in sd)


def _decoder_layout(cfg: CodecConfig) -> tp.Dict[int, str]:
    """Index in `SEANetDecoder.model` (seanet.py:409-479) -> streaming module name."""
    names: tp.Dict[int, str] = {0: "conv_pre_pointwise", 1: "conv_pre_depthwise"}
    idx = 2
    for i in range(len(cfg.strides)):
        idx += 2  # Scale (Identity at stage 0) + activation
        names[idx] = f"upsample_depthwise.{i}"
        names[idx + 1] = f"upsample_pointwise.{i}"
        idx += 2
        for j in range(cfg.n_residual_dec):
            names[idx] = f"blocks.{i}.{j}"
            idx += 1
    idx += 2  # Scale + activation
    names[idx] = "conv_post"
    return names


def training_to_streaming(sd: tp.Mapping[str, tp.Any], cfg: CodecConfig) -> "OrderedDict[str, Tensor]":
    """Rename the keys of a training-graph state dict to the streaming modules' names (no arithmetic)."""
    out: "OrderedDict[str, Tensor]" = OrderedDict()
    dec = _decoder_layout(cfg)
    unknown: tp.List[str] = []
    for k, v in sd.items():
        if k.startswith("module."):  # DistributedDataParallel prefix
            k = k[len("module."):]
        new: tp.Optional[str] = None
        if k.endswith(("_extra_state", "ema_embed", "ema_num", "spec.weight")) or k.startswith("disc"):
            continue
        m = re.fullmatch(r"quantizer\.layers\.(\d+)\.embed", k)
        if m:
            new = k
        elif k.endswith(("res_scale_param", "scale_param")) and k.startswith("encoder."):
            new = k
        elif (m := re.fullmatch(rf"encoder\.conv_pre\.1\.conv\.conv\.{_PARAM}", k)):
            new = f"encoder.conv_pre.{m[1]}"
        elif (m := re.fullmatch(rf"encoder\.blocks\.(\d+)\.(\d+)\.block\.(\d)\.conv\.conv\.{_PARAM}", k)):
            if int(m[3]) in _BLOCK_SLOT:
                new = f"encoder.blocks.{m[1]}.{m[2]}.{_BLOCK_SLOT[int(m[3])]}.{m[4]}"
        elif (m := re.fullmatch(rf"encoder\.(spec_blocks\.\d+|spec_post)\.layer\.conv\.conv\.{_PARAM}", k)):
            new = f"encoder.{m[1]}.layer.{m[2]}"
        elif (m := re.fullmatch(rf"encoder\.downsample\.(\d+)\.([23])\.conv\.conv\.{_PARAM}", k)):
            new = (f"encoder.downsample_pointwise.{m[1]}.1.{m[3]}" if m[2] == "2"
                   else f"encoder.downsample_depthwise.{m[1]}.{m[3]}")
        elif (m := re.fullmatch(rf"encoder\.conv_post\.([12])\.conv\.conv\.{_PARAM}", k)):
            new = f"encoder.conv_post_{'depthwise' if m[1] == '1' else 'pointwise'}.{m[2]}"
        elif (m := re.fullmatch(r"decoder\.model\.(\d+)\.res_scale_param", k)):
            if int(m[1]) in dec:
                new = f"decoder.{dec[int(m[1])]}.res_scale_param"
        elif (m := re.fullmatch(rf"decoder\.model\.(\d+)\.block\.(\d)\.conv\.conv\.{_PARAM}", k)):
            if int(m[1]) in dec and int(m[2]) in _BLOCK_SLOT:
                new = f"decoder.{dec[int(m[1])]}.{_BLOCK_SLOT[int(m[2])]}.{m[3]}"
        elif (m := re.fullmatch(rf"decoder\.model\.(\d+)\.(?:conv\.conv|convtr\.convtr)\.{_PARAM}", k)):
            if int(m[1]) in dec:
                new = f"decoder.{dec[int(m[1])]}.{m[2]}"
        if new is None:
            unknown.append(k)
            continue
        out[new] = torch.as_tensor(v).detach().cpu()
    if unknown:
        raise KeyError(f"unexpected keys in training state dict: {unknown[:8]}{' ...' if len(unknown) > 8 else ''}")
    return out


def deployment_weights(sd: tp.Mapping[str, tp.Any], cfg: CodecConfig, graph: str = "deploy",
                       norm: tp.Optional[str] = None,
                       norm_kwargs: tp.Optional[tp.Mapping[str, tp.Any]] = None) -> "OrderedDict[str, np.ndarray]":
    """Training-graph OR streaming state dict (reparametrised or not) -> the folded fp32 tensors the kernels
    consume, in `weights.tensor_shapes(cfg)` naming.  `graph="train"` keeps the training graph's output scaling
    (see fold.py; applied to an already-folded deployment dict as well); the decoder ResBlock `pre_scale` difference
    is a property of the graph, not of the weights.  `norm` / `norm_kwargs`: the training model's constructor
    arguments ("weight_norm" default, or "weight_standardization" with its dim / eps), see `fold.fold_state_dict`."""
    if is_training_state_dict(sd):
        sd = training_to_streaming(sd, cfg)
    folded = fold.fold_state_dict(OrderedDict(sd), cfg, part="", graph=graph, norm=norm, norm_kwargs=norm_kwargs)
    shapes = tensor_shapes(cfg)
    out: "OrderedDict[str, np.ndarray]" = OrderedDict()
    missing = []
    for name, shape in shapes.items():
        if name.endswith("spec.weight"):
            out[name] = dft_basis(shape[2])  # fixed buffer (causal_layers.py:115-129), not stored when not learnable
            continue
        if name not in folded:
            missing.append(name)
            continue
        t = folded[name].float().contiguous().numpy()
        if tuple(t.shape) != tuple(shape):
            raise ValueError(f"{name}: shape {tuple(t.shape)} != {tuple(shape)}")
        out[name] = t
    if missing:
        raise KeyError(f"missing tensors after conversion: {missing[:8]}{' ...' if len(missing) > 8 else ''}")
    return out


def load_checkpoint(path: str, cfg: CodecConfig, graph: str = "deploy", norm: tp.Optional[str] = None,
                    norm_kwargs: tp.Optional[tp.Mapping[str, tp.Any]] = None,
                    allow_pickle: bool = False) -> "OrderedDict[str, np.ndarray]":
    """Read a training checkpoint (`wrapper.py:428-444`: {"model": state_dict, "disc": ..., "epoch": ...}) or a
    bare state dict saved with torch.save, and convert it with `deployment_weights`.  Only tensors are needed, so
    the file is read with `weights_only=True` (no arbitrary pickle code from a third-party checkpoint runs);
    `allow_pickle=True` is the explicit opt-in for legacy files that hold other Python objects."""
    ckpt = torch.load(path, map_location="cpu", weights_only=not allow_pickle)
    sd = ckpt["model"] if isinstance(ckpt, dict) and "model" in ckpt and isinstance(ckpt["model"], dict) else ckpt
    return deployment_weights(sd, cfg, graph, norm, norm_kwargs)


# --------------------------------------------------------------------------------------- test support
def random_training_state_dict(cfg: CodecConfig, seed: int = 0) -> "OrderedDict[str, Tensor]":
    """A seeded training-format (`models.HILCodec.state_dict()`) checkpoint: weight-norm `weight_g/weight_v`
    pairs, non-zero `res_scale_param` / `scale_param`, codebooks.  Built by UN-folding
    `weights.random_weights(cfg, seed)` so activations keep O(1) magnitude; reproducible anywhere (numpy PCG64),
    which lets the GPU box rebuild the exact weights behind the committed training-graph fixtures."""
    rng = np.random.Generator(np.random.PCG64(seed + 7919))
    dep = random_weights(cfg, seed)
    out: "OrderedDict[str, Tensor]" = OrderedDict()

    def put(prefix: str, w: np.ndarray, b: tp.Optional[np.ndarray]) -> None:
        w = w.astype(np.float64)
        norm = np.sqrt((w.reshape(w.shape[0], -1) ** 2).sum(1)).reshape(-1, *([1] * (w.ndim - 1)))
        stretch = rng.uniform(0.5, 2.0, size=norm.shape)  # v is NOT the folded weight: g * v / ||v|| is
        out[prefix + "weight_g"] = torch.from_numpy(norm.astype(np.float32))
        out[prefix + "weight_v"] = torch.from_numpy((w * stretch).astype(np.float32))
        if b is not None:
            out[prefix + "bias"] = torch.from_numpy(b.astype(np.float32))

    def scalar(lo: float, hi: float) -> float:
        return float(rng.uniform(lo, hi))

    def res_block(src: str, dst: str, rs: float) -> None:
        p = scalar(0.7, 1.3)
        out[dst + "res_scale_param"] = torch.full((1,), p, dtype=torch.float32)
        for slot, name in _BLOCK_SLOT.items():
            w = dep[f"{src}{name}.weight"]
            b = dep.get(f"{src}{name}.bias")
            if name == "block.1.depthwise":  # ResBlock.merge_scaling folds res_scale * res_scale_param in here
                w, b = w / (rs * p), b / (rs * p)
            put(f"{dst}block.{slot}.conv.conv.", w, b)

    e = "encoder."
    put(e + "conv_pre.1.conv.conv.", dep[e + "conv_pre.weight"] * WAV_STD, dep[e + "conv_pre.bias"])
    n_stage = len(cfg.strides)
    for s in range(n_stage):
        for j in range(cfg.n_residual_enc):
            res_block(f"{e}blocks.{s}.{j}.", f"{e}blocks.{s}.{j}.", cfg.res_scale_enc)
    for s in range(n_stage + 1):
        name = f"spec_blocks.{s}" if s < n_stage else "spec_post"
        p = scalar(0.5, 1.0)
        out[f"{e}{name}.scale_param"] = torch.full((1,), p, dtype=torch.float32)
        std = fold.SPEC_STDS[min(s, 4)]
        put(f"{e}{name}.layer.conv.conv.", dep[f"{e}{name}.layer.weight"] * (std / (cfg.res_scale_enc * p)), None)
    for s in range(n_stage):
        put(f"{e}downsample.{s}.2.conv.conv.", dep[f"{e}downsample_pointwise.{s}.1.weight"], None)
        put(f"{e}downsample.{s}.3.conv.conv.", dep[f"{e}downsample_depthwise.{s}.weight"],
            dep[f"{e}downsample_depthwise.{s}.bias"])
    put(e + "conv_post.1.conv.conv.", dep[e + "conv_post_depthwise.weight"], None)
    put(e + "conv_post.2.conv.conv.", dep[e + "conv_post_pointwise.weight"], dep[e + "conv_post_pointwise.bias"])

    d = "decoder."
    for idx, name in _decoder_layout(cfg).items():
        t = f"{d}model.{idx}."
        if name.startswith("blocks."):
            res_block(f"{d}{name}.", t, cfg.res_scale_dec)
        elif name.startswith("upsample_depthwise"):
            put(t + "convtr.convtr.", dep[f"{d}{name}.weight"], None)
        elif name == "conv_post":
            put(t + "conv.conv.", dep[d + "conv_post.weight"] / WAV_STD, dep[d + "conv_post.bias"])
        else:
            put(t + "conv.conv.", dep[f"{d}{name}.weight"], dep.get(f"{d}{name}.bias"))
    for i in range(cfg.num_quantizers):
        out[f"quantizer.layers.{i}.embed"] = torch.from_numpy(dep[f"quantizer.layers.{i}.embed"].copy())
    return out
