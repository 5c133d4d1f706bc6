"""Import the reference's own deployment classes (dev container only).

TEST INFRASTRUCTURE ONLY.  `/root/reference` exists in the dev container and not on
the GPU box, so this is used by `tests/golden/make_golden.py` (fixture generation)
and by the `not gpu` tests that pin the oracle against the reference; it is skipped
when the reference tree is absent.

`import models.hilcodec` runs the training wrapper's imports (librosa, pesq, ...),
which are not installed; the deployment classes themselves only need torch.  So the
package objects `models` / `models.hilcodec` are pre-registered with the right
`__path__` (their `__init__` is never executed) and `librosa.filters.mel` is stubbed
(functional/audio_functional.py:8 imports it at module scope; never called here).
"""
from __future__ import annotations

import os
import sys
import types
import warnings
from typing import Dict

_FULL = os.environ.get("HILCODEC_REFERENCE", "/root/reference")
_COPY = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")   # oracle/build_ref.py


def available() -> bool:
    """The full reference tree (dev container): training graph, generic modules, ONNX files."""
    return os.path.isfile(os.path.join(_FULL, "models", "hilcodec", "streaming.py")) and \
        os.path.isdir(os.path.join(_FULL, "onnx"))


def deploy_available() -> bool:
    """The deployment classes are importable: full tree, or the unmodified copy in oracle/_ref (GPU box)."""
    return available() or os.path.isfile(os.path.join(_COPY, "models", "hilcodec", "streaming.py"))


REF = _FULL if available() or not deploy_available() else _COPY


def _stub_unused_imports() -> None:
    """streaming.py:18-19 imports `functional.STDCT` and `utils.verbose` at module scope; neither is used by the
    deployment forward (verbose() only guards a constructor warning).  In the oracle/_ref copy those packages are
    absent: register stand-ins so the unmodified file imports."""
    if REF == _COPY:
        if "functional" not in sys.modules:
            f = types.ModuleType("functional")
            f.STDCT = f.STFT = None
            sys.modules["functional"] = f
        if "utils" not in sys.modules:
            u = types.ModuleType("utils")
            u.verbose = lambda: False
            sys.modules["utils"] = u


def import_streaming():
    """Return the reference module `models.hilcodec.streaming`."""
    if not deploy_available():
        raise ImportError(f"reference tree not found at {_FULL} (nor a copy at {_COPY})")
    _stub_unused_imports()
    if "librosa" not in sys.modules:
        lib = types.ModuleType("librosa")
        filt = types.ModuleType("librosa.filters")
        filt.mel = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("librosa stub"))
        lib.filters = filt
        sys.modules["librosa"] = lib
        sys.modules["librosa.filters"] = filt
    for name, sub in (("models", "models"), ("models.hilcodec", os.path.join("models", "hilcodec"))):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = [os.path.join(REF, sub)]
            sys.modules[name] = m
    if REF not in sys.path:
        sys.path.insert(0, REF)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from models.hilcodec import streaming  # type: ignore
    return streaming


def build_reference_model(weights: Dict[str, "object"], num_quantizers: int):
    """Reference `streaming.HILCodec` (eval, folded) loaded with deployment weights.

    Follows SURVEY.md appendix B: construct with the yaml `model_kwargs` minus
    `spec_learnable/causal/pad_mode`, fold, then `load_state_dict(strict=False)`."""
    import torch
    import yaml

    streaming = import_streaming()
    with open(os.path.join(REF, "configs", "hilcodec_music.yaml")) as f:
        kw = yaml.safe_load(f)["model_kwargs"]
    for k in ("spec_learnable", "causal", "pad_mode"):
        kw.pop(k)
    kw["vq_kwargs"]["num_quantizers"] = num_quantizers
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = streaming.HILCodec(24000, **kw).eval()
        model.remove_weight_reparameterizations()
    enc = {k[len("encoder."):]: torch.as_tensor(v) for k, v in weights.items() if k.startswith("encoder.")}
    dec = {k[len("decoder."):]: torch.as_tensor(v) for k, v in weights.items() if k.startswith("decoder.")}
    r = model.encoder.load_state_dict(enc, strict=False)
    assert not r.unexpected_keys, r.unexpected_keys
    assert all("scale_param" in k for k in r.missing_keys), r.missing_keys
    r = model.decoder.load_state_dict(dec, strict=False)
    assert not r.unexpected_keys, r.unexpected_keys
    assert all("scale_param" in k for k in r.missing_keys), r.missing_keys
    for i in range(num_quantizers):
        e = torch.as_tensor(weights[f"quantizer.layers.{i}.embed"])
        model.quantizer.layers[i].embed.copy_(e)
        model.dequantizer.layers[i].embed.copy_(e)
    return model


def reference_forward(model, x, n, enc_caches=None, dec_caches=None):
    """The four-call flow of scripts/HILCodec Onnx.ipynb cell 3."""
    import torch

    with torch.no_grad():
        if enc_caches is None or dec_caches is None:
            ce, cd = model.initialize_cache(x)
            enc_caches = enc_caches or ce
            dec_caches = dec_caches or cd
        z, ce = model.encoder(x, *enc_caches)
        idx = model.quantizer(z, n)
        q = model.dequantizer(idx, n)
        y, cd = model.decoder(q, *dec_caches)
    return {"z": z, "indices": idx, "q": q, "wav": y, "enc_caches": list(ce), "dec_caches": list(cd)}


# ------------------------------------------------------------------ training graph
def _model_kwargs(num_quantizers: int) -> dict:
    import yaml

    with open(os.path.join(REF, "configs", "hilcodec_music.yaml")) as f:
        kw = yaml.safe_load(f)["model_kwargs"]
    kw["vq_kwargs"]["num_quantizers"] = num_quantizers
    return kw


def build_reference_training_model(state_dict, num_quantizers: int, norm: str = "weight_norm", norm_kwargs=None):
    """Reference TRAINING-graph `models.hilcodec.models.HILCodec` (eval) loaded with a training-format state
    dict (`hilcodec_b200.checkpoint.random_training_state_dict`).  The EMA statistics and `_extra_state` of the
    codebooks keep their constructor values (they do not enter the eval forward).  `state_dict=None` keeps the
    constructor's random initialisation (used for the weight-standardisation parametrisation, whose keys the
    reference writes itself)."""
    import torch

    import_streaming()  # registers the package stubs
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from models.hilcodec.models import HILCodec  # type: ignore
        model = HILCodec(24000, **{**_model_kwargs(num_quantizers), "norm": norm, "norm_kwargs": dict(norm_kwargs or {})}).eval()
    if state_dict is None:
        for layer in model.quantizer.layers:
            layer.initted = True
        return model
    r = model.load_state_dict({k: torch.as_tensor(v) for k, v in state_dict.items()}, strict=False)
    assert not r.unexpected_keys, r.unexpected_keys
    assert all(k.endswith(("ema_embed", "ema_num", "_extra_state", "spec.weight")) for k in r.missing_keys), r.missing_keys
    for layer in model.quantizer.layers:
        layer.initted = True  # skip the k-means initialisation branch (vector_quantize.py:138-139)
    return model


def streaming_model_from_training(train_model, num_quantizers: int):
    """`scripts/HILCodec Onnx.ipynb` cells 0-1, restated: build the reference `streaming.HILCodec`, copy every
    conv of the training model into it, fold.  Returns the folded streaming model (the deployment weights the
    reference itself would export)."""
    streaming = import_streaming()
    kw = _model_kwargs(num_quantizers)
    for k in ("spec_learnable", "causal", "pad_mode"):
        kw.pop(k)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = streaming.HILCodec(24000, **kw).eval()
    me, enc = model.encoder, train_model.encoder
    me.conv_pre.load_state_dict(enc.conv_pre[1].conv.conv.state_dict())
    for mine, theirs in zip(me.blocks, enc.blocks):
        for a, b in zip(mine, theirs):
            a.block[0].pointwise[1].load_state_dict(b.block[1].conv.conv.state_dict())
            a.block[0].depthwise.load_state_dict(b.block[2].conv.conv.state_dict())
            a.block[1].pointwise[1].load_state_dict(b.block[4].conv.conv.state_dict())
            a.block[1].depthwise.load_state_dict(b.block[5].conv.conv.state_dict())
            a.res_scale_param.data.copy_(b.res_scale_param.data)
    for a, b in zip(me.spec_blocks, enc.spec_blocks):
        a.layer.load_state_dict(b.layer.conv.conv.state_dict())
        a.scale_param.data.copy_(b.scale_param.data)
    for p, d, b in zip(me.downsample_pointwise, me.downsample_depthwise, enc.downsample):
        p[1].load_state_dict(b[2].conv.conv.state_dict())
        d.load_state_dict(b[3].conv.conv.state_dict())
    me.spec_post.layer.load_state_dict(enc.spec_post.layer.conv.conv.state_dict())
    me.spec_post.scale_param.data.copy_(enc.spec_post.scale_param.data)
    me.conv_post_depthwise.load_state_dict(enc.conv_post[1].conv.conv.state_dict())
    me.conv_post_pointwise.load_state_dict(enc.conv_post[2].conv.conv.state_dict())
    md, dec = model.decoder, train_model.decoder.model
    md.conv_pre_pointwise.load_state_dict(dec[0].conv.conv.state_dict())
    md.conv_pre_depthwise.load_state_dict(dec[1].conv.conv.state_dict())
    idx = 2
    for ud, up, blocks in zip(md.upsample_depthwise, md.upsample_pointwise, md.blocks):
        idx += 2
        ud.load_state_dict(dec[idx].convtr.convtr.state_dict()); idx += 1
        up.load_state_dict(dec[idx].conv.conv.state_dict()); idx += 1
        for blk in blocks:
            blk.block[0].pointwise[1].load_state_dict(dec[idx].block[1].conv.conv.state_dict())
            blk.block[0].depthwise.load_state_dict(dec[idx].block[2].conv.conv.state_dict())
            blk.block[1].pointwise[1].load_state_dict(dec[idx].block[4].conv.conv.state_dict())
            blk.block[1].depthwise.load_state_dict(dec[idx].block[5].conv.conv.state_dict())
            blk.res_scale_param.data.copy_(dec[idx].res_scale_param.data)
            idx += 1
    idx += 2
    md.conv_post.load_state_dict(dec[idx].conv.conv.state_dict())
    for a, b, c in zip(model.quantizer.layers, model.dequantizer.layers, train_model.quantizer.layers):
        a.embed.data.copy_(c.embed.data)
        b.embed.data.copy_(c.embed.data)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model.remove_weight_reparameterizations()
    return model


def reference_training_forward(train_model, x, n):
    """`models.HILCodec.forward` (models.py:111-118), spelled out to also return the latents and indices."""
    import torch

    with torch.no_grad():
        z = train_model.encoder(x)
        q, num_replaces, loss_vq, idx = train_model.quantizer(z, n, return_indices=True)
        y = train_model.decoder(q).float()
        y2, nr2, loss2 = train_model(x, n)
    assert torch.equal(y, y2) and torch.equal(loss_vq, loss2)
    return {"z": z, "q": q, "indices": idx, "loss_vq": loss_vq, "wav": y, "num_replaces": num_replaces}
