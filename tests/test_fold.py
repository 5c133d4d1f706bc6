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
