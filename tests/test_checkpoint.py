"""SURVEY.md 8f.2 / 8f.3 on the CPU: training-checkpoint conversion (hilcodec_b200/checkpoint.py) and the
training-graph restatement of the oracle, pinned against the reference's own `models.HILCodec`
(dev container) and against the committed fixture made from it (anywhere)."""
import os
import re

import numpy as np
import pytest
import torch

from hilcodec_b200 import checkpoint, fold
from hilcodec_b200 import weights as W
from oracle import hilcodec_oracle as O
from oracle import ref_shim

from helpers import GOLDEN, params

need_ref = pytest.mark.skipif(not ref_shim.available(), reason="reference tree absent (GPU box)")


def _fixture():
    return np.load(os.path.join(GOLDEN, "ref_train_random.npz"))


def test_training_keys_rename_to_streaming_keys():
    cfg = W.CodecConfig(num_quantizers=3)
    sd = checkpoint.random_training_state_dict(cfg, 0)
    assert checkpoint.is_training_state_dict(sd)
    st = checkpoint.training_to_streaming(sd, cfg)
    assert not checkpoint.is_training_state_dict(st)
    # one renamed key per training key (embed / scale params keep their names)
    assert len(st) == len(sd)
    assert st["decoder.conv_post.weight_v"].shape == (1, 96, 5)
    assert st["decoder.upsample_depthwise.3.weight_v"].shape == (192, 1, 4)
    assert st["decoder.blocks.0.2.block.1.depthwise.bias"].shape == (768,)
    assert st["encoder.downsample_pointwise.2.1.weight_g"].shape == (512, 1, 1)
    # DDP prefix and training-only buffers are tolerated, anything else is an error
    extra = {"module." + k: v for k, v in sd.items()}
    extra["module.quantizer.layers.0.ema_num"] = torch.zeros(1024)
    extra["module.quantizer.layers.0._extra_state"] = {"initted": True}
    assert list(checkpoint.training_to_streaming(extra, cfg)) == list(st)
    with pytest.raises(KeyError):
        checkpoint.training_to_streaming({**sd, "decoder.model.99.conv.conv.bias": torch.zeros(1)}, cfg)


def test_unfolded_random_checkpoint_folds_back_to_random_weights():
    """random_training_state_dict() un-folds weights.random_weights(); conversion must undo it."""
    cfg = W.CodecConfig(num_quantizers=2)
    dep = W.random_weights(cfg, 5)
    got = checkpoint.deployment_weights(checkpoint.random_training_state_dict(cfg, 5), cfg)
    assert list(got) == list(W.tensor_shapes(cfg))
    for k, v in got.items():
        if re.search(r"spec.*layer\.bias", k):
            continue  # the training graph has no such bias: it is created by folding mean/std
        assert np.allclose(v, dep[k], rtol=2e-6, atol=1e-7), k
    tr = checkpoint.deployment_weights(checkpoint.random_training_state_dict(cfg, 5), cfg, graph="train")
    for k in got:
        if k == "decoder.conv_post.bias":
            assert np.allclose(tr[k], got[k] * np.float32(W.WAV_STD))
        else:
            assert np.array_equal(tr[k], got[k]), k
    assert np.array_equal(fold.to_train_graph(got)["decoder.conv_post.bias"], tr["decoder.conv_post.bias"])


def test_load_checkpoint_reads_wrapper_format(tmp_path):
    cfg = W.CodecConfig(num_quantizers=2)
    sd = checkpoint.random_training_state_dict(cfg, 1)
    path = os.path.join(str(tmp_path), "00010.pth")
    torch.save({"model": sd, "disc": {}, "epoch": 10}, path)  # wrapper.py:428-444
    a = checkpoint.load_checkpoint(path, cfg)
    b = checkpoint.deployment_weights(sd, cfg)
    assert all(np.array_equal(a[k], b[k]) for k in b)
    torch.save(sd, path)
    a = checkpoint.load_checkpoint(path, cfg, graph="train")
    assert np.allclose(a["decoder.conv_post.bias"], b["decoder.conv_post.bias"] * W.WAV_STD)


def test_oracle_training_graph_matches_fixture():
    """The committed outputs of the reference's TRAINING graph (tests/golden/make_golden.py train)."""
    g = _fixture()
    n_q, n = int(g["n_q"]), int(g["n"])
    cfg = W.CodecConfig(num_quantizers=n_q)
    p = params(checkpoint.deployment_weights(checkpoint.random_training_state_dict(cfg, int(g["seed"])), cfg, "train"))
    ocfg = O.CodecConfig(num_quantizers=n_q)
    for T in g["lengths"].tolist():
        x = torch.from_numpy(g[f"x_{T}"])
        with torch.no_grad():
            o = O.codec_forward_train(ocfg, p, x, n)
        frames = -(-T // 320)
        assert o["z"].shape == (x.shape[0], 128, frames) and o["wav"].shape == (x.shape[0], 1, 320 * frames)
        assert np.abs(o["z"].numpy() - g[f"z_{T}"]).max() < 2e-5
        assert np.array_equal(o["indices"].numpy(), g[f"indices_{T}"].astype(np.int64))
        assert np.array_equal(o["q"].numpy(), g[f"q_{T}"])
        assert np.abs(o["wav"].numpy() - g[f"wav_{T}"]).max() < 2e-5
        assert abs(float(o["loss_vq"]) - float(g[f"loss_{T}"])) < 1e-5


def test_training_graph_differs_from_deploy_graph_only_where_documented():
    """Quirks 1, 2 and 4: same encoder arithmetic and (away from ties) the same indices on hop multiples;
    the decoders differ."""
    cfg = W.CodecConfig(num_quantizers=4)
    dep = params(W.random_weights(cfg, 9))
    tr = params(fold.to_train_graph(W.random_weights(cfg, 9)))
    ocfg = O.CodecConfig(num_quantizers=4)
    x = (0.1 * torch.randn(2, 1, 320 * 6, generator=torch.Generator().manual_seed(0))).clamp(-1, 1)
    with torch.no_grad():
        a = O.codec_forward(ocfg, dep, x, 4)
        b = O.codec_forward_train(ocfg, tr, x, 4)
    assert (a["z"].transpose(1, 2) - b["z"]).abs().max().item() < 1e-5
    assert torch.equal(a["indices"].permute(1, 0, 2), b["indices"])
    assert (a["wav"] - b["wav"]).abs().max().item() > 1e-3


@need_ref
def test_conversion_matches_reference_notebook_flow():
    """`scripts/HILCodec Onnx.ipynb` cell 1 + remove_weight_reparameterizations(), run on the reference's
    own classes, against checkpoint.deployment_weights()."""
    cfg = W.CodecConfig(num_quantizers=3)
    sd = checkpoint.random_training_state_dict(cfg, 2)
    train_model = ref_shim.build_reference_training_model(sd, 3)  # asserts the key naming (strict load)
    ref = ref_shim.streaming_model_from_training(train_model, 3)
    ref_w = {**{"encoder." + k: v for k, v in ref.encoder.state_dict().items()},
             **{"decoder." + k: v for k, v in ref.decoder.state_dict().items()}}
    mine = checkpoint.deployment_weights(sd, cfg)
    for k, v in mine.items():
        if k.startswith("quantizer."):
            assert np.array_equal(v, ref.quantizer.layers[int(k.split(".")[2])].embed.numpy())
        else:
            r = ref_w[k].numpy()
            assert np.abs(v - r).max() <= 1e-6 * max(1.0, np.abs(r).max()), k


@need_ref
@pytest.mark.parametrize("T", [1, 319, 320, 321, 1000, 2560])
def test_oracle_training_graph_matches_reference_live(T):
    cfg = W.CodecConfig(num_quantizers=3)
    sd = checkpoint.random_training_state_dict(cfg, 4)
    model = ref_shim.build_reference_training_model(sd, 3)
    p = params(checkpoint.deployment_weights(sd, cfg, "train"))
    x = (0.1 * torch.randn(2, 1, T, generator=torch.Generator().manual_seed(T))).clamp(-1, 1)
    r = ref_shim.reference_training_forward(model, x, 2)
    with torch.no_grad():
        o = O.codec_forward_train(O.CodecConfig(num_quantizers=3), p, x, 2)
    assert o["z"].shape == r["z"].shape and o["wav"].shape == r["wav"].shape
    assert (o["z"] - r["z"]).abs().max().item() < 2e-5
    assert torch.equal(o["indices"], r["indices"])
    assert (o["wav"] - r["wav"]).abs().max().item() < 2e-5
    assert abs(float(o["loss_vq"]) - float(r["loss_vq"])) < 1e-5
