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

REF = os.environ.get("HILCODEC_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "models", "hilcodec", "streaming.py"))


def import_streaming():
    """Return the reference module `models.hilcodec.streaming`."""
    if not available():
        raise ImportError(f"reference tree not found at {REF}")
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
