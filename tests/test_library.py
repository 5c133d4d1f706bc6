"""CPU-side checks of the boundary: the shared library builds, loads and exports every
symbol include/hilcodec_b200.h declares; host-side argument validation mirrors the
reference's error behaviour.  No compute calls (no GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from hilcodec_b200 import _lib, build
from hilcodec_b200 import streaming as S
from hilcodec_b200 import weights as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_header_symbols():
    lib_path = build.build_library()
    assert os.path.exists(lib_path)
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "hilcodec_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(hil_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.hil_abi_version() == 1


def test_config_struct_layout_matches_header_defaults():
    lib = _lib.load()
    c = _lib.HilConfig()
    lib.hil_config_default(C.byref(c), 12)
    assert (c.channels_enc, c.channels_dec, c.n_fft_base) == (64, 96, 64)
    assert (c.n_residual_enc, c.n_residual_dec) == (2, 3)
    assert c.res_scale_enc == 0.5773502691896258 and c.res_scale_dec == 0.5773502691896258
    assert list(c.strides)[:4] == [8, 5, 4, 2] and c.n_strides == 4
    assert (c.kernel_size, c.dim, c.codebook_size, c.num_quantizers) == (5, 128, 1024, 12)


def test_model_create_validates_config_without_gpu():
    lib = _lib.load()
    c = _lib.HilConfig()
    lib.hil_config_default(C.byref(c), 8)
    c.kernel_size = 7
    h = C.c_void_p()
    assert lib.hil_model_create(C.byref(c), C.byref(h)) == -1
    assert b"kernel_size" in lib.hil_last_error()
    lib.hil_config_default(C.byref(c), 8)
    c.n_residual_dec = 99
    assert lib.hil_model_create(C.byref(c), C.byref(h)) == -1 and b"n_residual" in lib.hil_last_error()
    lib.hil_config_default(C.byref(c), 8)
    assert lib.hil_model_create(C.byref(c), C.byref(h)) == 0
    # finalize with nothing set: HIL_ERR_MISSING, no CUDA call made
    assert lib.hil_model_finalize(h) == -2
    lib.hil_model_destroy(h)


def test_dropin_surface_and_errors():
    m = S.HILCodec(24000, vq_kwargs=dict(dim=128, codebook_size=1024, num_quantizers=8))
    w = W.random_weights(W.HIL_SPEECH, 3)
    r = m.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()})
    assert not r.missing_keys and not r.unexpected_keys
    assert m.encoder.num_cache == 22 and m.decoder.num_cache == 30
    assert m.encoder.hop_length == 320 and m.encoder.ratios == [2, 4, 5, 8]
    assert len(m.quantizer.layers) == 8
    assert m.quantizer.layers[0].embed.shape == (1024, 128)
    sd = m.state_dict()
    for k in ("encoder.conv_pre.weight", "encoder.blocks.0.0.block.0.pointwise.1.weight",
              "decoder.upsample_depthwise.3.weight", "quantizer.layers.7.embed", "dequantizer.layers.7.embed"):
        assert k in sd
    assert set(m.encoder.state_dict()) == {k[len("encoder."):] for k in W.tensor_shapes(W.HIL_SPEECH) if k.startswith("encoder.")}
    # no CPU path: CPU tensors raise instead of silently computing somewhere else
    with pytest.raises(RuntimeError, match="CUDA"):
        m.encoder(torch.zeros(1, 1, 320))
    with pytest.raises(RuntimeError, match="CUDA"):
        m.quantizer(torch.zeros(1, 1, 128), 8)
    with pytest.raises(ValueError, match="Unknown norm"):
        S.HILCodec(24000, norm="layer_norm")
    with pytest.raises(RuntimeError, match="missing"):
        S.HILCodec(24000, vq_kwargs=dict(dim=128, num_quantizers=8)).load_state_dict({})


def test_graph_selection_without_gpu():
    """hil_model_set_graph: deploy (default) / train before finalize, rejected values, NULL handles."""
    lib = _lib.load()
    c = _lib.HilConfig()
    lib.hil_config_default(C.byref(c), 8)
    h = C.c_void_p()
    assert lib.hil_model_create(C.byref(c), C.byref(h)) == 0
    assert lib.hil_model_graph(h) == _lib.HIL_GRAPH_DEPLOY
    assert lib.hil_model_set_graph(h, _lib.HIL_GRAPH_TRAIN) == 0 and lib.hil_model_graph(h) == _lib.HIL_GRAPH_TRAIN
    assert lib.hil_model_set_graph(h, 7) == -1 and b"unknown graph" in lib.hil_last_error()
    assert lib.hil_model_graph(h) == _lib.HIL_GRAPH_TRAIN
    assert lib.hil_model_set_graph(None, 0) == -1 and lib.hil_model_graph(None) == -1
    # calls on a model that was never finalized fail before any CUDA call
    assert lib.hil_encode_ragged(h, None, None, 1, 100, None, None) == -1
    lib.hil_model_destroy(h)


def test_training_graph_adapter_surface_without_gpu():
    """hilcodec_b200/models.py: constructors, graph plumbing and the no-CPU-path rule."""
    from hilcodec_b200 import checkpoint, models
    cfg = W.CodecConfig(num_quantizers=3)
    m = models.HILCodec.from_training_state_dict(checkpoint.random_training_state_dict(cfg, 0), 3)
    assert m.graph == "train" and m.deploy._core.graph == _lib.HIL_GRAPH_TRAIN and m.quantizer.num_quantizers == 3
    d = models.HILCodec.from_training_state_dict(checkpoint.random_training_state_dict(cfg, 0), 3, graph="deploy")
    assert d.deploy._core.graph == _lib.HIL_GRAPH_DEPLOY
    bias_t = m.deploy._core.weights["decoder.conv_post.bias"]
    bias_d = d.deploy._core.weights["decoder.conv_post.bias"]
    assert torch.allclose(bias_t, bias_d * W.WAV_STD)
    t2 = models.HILCodec(d.deploy, graph="train")          # same weights, training-graph output scaling
    assert t2.deploy is not d.deploy and t2.deploy._core.graph == _lib.HIL_GRAPH_TRAIN
    assert torch.allclose(t2.deploy._core.weights["decoder.conv_post.bias"], bias_t)
    with pytest.raises(ValueError, match="Unknown graph"):
        models.HILCodec(d.deploy, graph="onnx")
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 1, 333))


def test_onnx_runner_report_and_cache_names():
    """Host logic of the test_onnx.py-compatible runner: the timing lines (test_onnx.py:41-47) and the cache tensor
    names of the exported graphs (`e_in0..21`, `d_in0..29`), which the reference's own cache files carry."""
    import numpy as np

    from hilcodec_b200 import onnx_runner
    from oracle import ref_shim

    text = onnx_runner.report(24000 * 5, 24000, 3.2, 9.0)
    assert "wav length: 5.0 s" in text
    assert "encoder: 3.2 s / rtf: 1.5625 (↑)" in text and "decoder: 9.0 s / rtf: 0.5556 (↑)" in text
    assert "decoder" not in onnx_runner.report(24000, 24000, 1.0, 0.0)
    assert onnx_runner.cache_names("enc", 3) == ["e_in0", "e_in1", "e_in2"]
    assert onnx_runner.cache_names("dec", 2) == ["d_in0", "d_in1"]
    if ref_shim.available():
        for side, n in (("enc", 22), ("dec", 30)):
            ref = np.load(os.path.join(ref_shim.REF, "onnx", f"hil_music_cache_{side}.npz"))
            assert sorted(ref.files, key=lambda k: int(k.split("in")[1])) == onnx_runner.cache_names(side, n)
