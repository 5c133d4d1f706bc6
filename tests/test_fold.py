"""Load-time weight folding (hilcodec_b200/fold.py) against the reference's own
`remove_weight_reparameterizations` (streaming.py:740-747).  Dev container only."""
import warnings

import pytest
import torch

from hilcodec_b200 import fold
from hilcodec_b200 import weights as W
from oracle import ref_shim


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree absent (GPU box)")
def test_fold_matches_reference_merge():
    import yaml, os
    streaming = ref_shim.import_streaming()
    with open(os.path.join(ref_shim.REF, "configs", "hilcodec_speech.yaml")) as f:
        kw = yaml.safe_load(f)["model_kwargs"]
    for k in ("spec_learnable", "causal", "pad_mode"):
        kw.pop(k)
    torch.manual_seed(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = streaming.HILCodec(24000, **kw).eval()
    with torch.no_grad():  # zero_init leaves the scale params at 0; make them matter
        for n, p in model.named_parameters():
            if n.endswith("scale_param"):
                p.fill_(0.3 + 0.01 * (hash(n) % 17))
    raw = {k: v.clone() for k, v in model.state_dict().items()}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model.remove_weight_reparameterizations()
    ref = model.state_dict()
    mine = fold.fold_state_dict(raw, W.HIL_SPEECH, part="")
    shapes = W.tensor_shapes(W.HIL_SPEECH)
    checked = 0
    for k in shapes:
        if k.startswith("quantizer."):
            continue
        assert k in mine, k
        assert torch.allclose(mine[k].float(), ref[k].float(), rtol=1e-6, atol=1e-7), k
        checked += 1
    assert checked == len([k for k in shapes if not k.startswith("quantizer.")])


def test_folded_dict_passes_through():
    w = {k: torch.from_numpy(v) for k, v in W.random_weights(W.HIL_SPEECH, 0).items()}
    out = fold.fold_state_dict(w, W.HIL_SPEECH)
    assert all(torch.equal(out[k], w[k]) for k in w)


# ------------------------------------------------------------------ weight standardisation
def _ws_reference(v, g, scale, dim, eps):
    """fp64 restatement of the published formula (modules/weight_standardization.py:112-128 docstring):
    weight = (gain * scale) * (v - mean(v)) / sqrt(var(v) * fan_in)."""
    v = v.double()
    axes = [a for a in range(v.dim()) if a not in dim]
    fan_in = 1
    for a in axes:
        fan_in *= v.shape[a]
    mean = v.mean(dim=axes, keepdim=True)
    var = ((v - mean) ** 2).mean(dim=axes, keepdim=True)
    w = (v - mean) / torch.sqrt(torch.clamp(var * fan_in, min=eps))
    if g is not None:
        w = w * (g.double() * (scale.double() if scale is not None else 1.0))
    return w


@pytest.mark.parametrize("shape,dim", [((8, 6, 1), (0,)), ((8, 1, 5), (0,)), ((6, 1, 4), (0,)), ((4, 3, 5), (0, 1))])
def test_standardize_weight_formula(shape, dim):
    g = torch.Generator().manual_seed(3)
    v = torch.randn(*shape, generator=g)
    gshape = [shape[a] if a in dim else 1 for a in range(len(shape))]
    gain = torch.rand(*gshape, generator=g) + 0.5
    scale = torch.tensor([0.7])
    for gg, ss in ((gain, scale), (gain, None), (None, None)):
        got = fold.standardize_weight(v, gg, ss, dim=dim if len(dim) > 1 else dim[0], eps=1e-7)
        ref = _ws_reference(v, gg, ss, dim, 1e-7)
        assert got.dtype == torch.float32
        assert (got.double() - ref).abs().max().item() < 1e-6
    if len(dim) == 1:  # every output row has zero mean and norm 1 (= 1/sqrt(fan_in) std)
        w = fold.standardize_weight(v, None, None, dim=0)
        assert w.reshape(shape[0], -1).sum(1).abs().max().item() < 1e-5
        assert (w.reshape(shape[0], -1).norm(dim=1) - 1).abs().max().item() < 1e-5


def test_detect_norm_from_keys():
    assert fold.detect_norm({"a.weight_v": 0, "a.weight_g": 0, "a.weight_scale": 0}) == "weight_standardization"
    assert fold.detect_norm({"a.weight_v": 0, "a.bias": 0}) == "weight_standardization"   # learnable_gain=False
    assert fold.detect_norm({"a.parametrizations.weight.original0": 0}) == "weight_norm"
    assert fold.detect_norm({"a.weight_v": 0, "a.weight_g": 0}) is None                   # both norms write these
    with pytest.raises(ValueError):
        fold.fold_state_dict({"a.weight_v": torch.zeros(2, 2, 1), "a.weight_g": torch.zeros(2, 1, 1)}, W.HIL_SPEECH,
                             norm="spectral_norm")


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree absent (GPU box)")
@pytest.mark.parametrize("kwargs", [dict(), dict(scale=0.5), dict(learnable_gain=False), dict(eps=1e-3, scale=2.0)])
def test_standardize_weight_matches_reference_module(kwargs):
    """`weight_standardization()` of the reference itself (modules/weight_standardization.py:112-134) on Conv1d and
    ConvTranspose1d: the weight its forward pre-hook computes (`compute_weight` :30-41) against fold.py.  (The
    reference's own `remove()` :100-106 cannot run -- it deletes `weight_scale` from `_parameters`, where a buffer
    never is -- so the hook's weight is the pin.)"""
    ref_shim.import_streaming()
    from modules.weight_standardization import weight_standardization  # type: ignore
    torch.manual_seed(1)
    for conv in (torch.nn.Conv1d(6, 8, 1), torch.nn.Conv1d(8, 8, 5, groups=8), torch.nn.ConvTranspose1d(8, 8, 4, 2, groups=8)):
        m = weight_standardization(conv, **kwargs)
        with torch.no_grad():
            if kwargs.get("learnable_gain", True):
                m.weight_g.uniform_(0.5, 1.5)
        sd = {"x." + k: v.clone() for k, v in m.state_dict().items()}
        assert ("x.weight_scale" in sd) == ("scale" in kwargs)
        assert fold.detect_norm(sd) in ("weight_standardization", None)
        mine = fold._remove_weight_standardization(sd, eps=kwargs.get("eps", 1e-7))
        with torch.no_grad():
            m(torch.zeros(1, m.in_channels, 8))   # the forward pre-hook recomputes m.weight
        assert set(mine) == {"x.weight"} | ({"x.bias"} if m.bias is not None else set())
        assert torch.equal(mine["x.weight"], m.weight.detach())


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree absent (GPU box)")
@pytest.mark.parametrize("norm_kwargs", [dict(), dict(scale=1.5, eps=1e-6)])
def test_weight_standardized_checkpoint_serves_like_the_reference(norm_kwargs):
    """A reference TRAINING model built with norm="weight_standardization" (models.py:49, conv.py:36-37): its own
    state_dict(), converted by checkpoint.deployment_weights(norm=...), must reproduce its forward through the
    oracle's training graph (latents, indices, PCM)."""
    from hilcodec_b200 import checkpoint
    from oracle import hilcodec_oracle as O
    from helpers import params

    torch.manual_seed(5)
    model = ref_shim.build_reference_training_model(None, 3, "weight_standardization", norm_kwargs)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith("scale_param"):
                p.fill_(0.4)
            elif n.endswith("weight_g"):
                p.uniform_(0.8, 1.2)
        for layer in model.quantizer.layers:
            layer.embed.copy_(torch.randn_like(layer.embed) * (11.3 / 128 ** 0.5))
    sd = model.state_dict()
    assert fold.detect_norm(sd) == ("weight_standardization" if "scale" in norm_kwargs else None)
    cfg = W.CodecConfig(num_quantizers=3)
    w = checkpoint.deployment_weights(sd, cfg, "train", norm="weight_standardization", norm_kwargs=norm_kwargs)
    x = (0.1 * torch.randn(2, 1, 1600, generator=torch.Generator().manual_seed(2))).clamp(-1, 1)
    r = ref_shim.reference_training_forward(model, x, 3)
    with torch.no_grad():
        o = O.codec_forward_train(O.CodecConfig(num_quantizers=3), params(w), x, 3)
    assert (o["z"] - r["z"]).abs().max().item() < 2e-5
    assert torch.equal(o["indices"], r["indices"])
    assert (o["wav"] - r["wav"]).abs().max().item() < 2e-5
