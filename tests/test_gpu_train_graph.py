"""SURVEY.md 8f.2 / 8f.3 on the GPU: a training-format checkpoint served through the TRAINING graph
(`hil_model_set_graph(HIL_GRAPH_TRAIN)`, `hil_encode_ragged`) against the committed outputs of the reference's
own `models/hilcodec/models.py` `HILCodec.forward` (tests/golden/ref_train_random.npz) and the oracle.

Bars as everywhere: latents and PCM max-abs-err < 1e-4; indices equal, a disagreement being tolerated only at a
near tie (fp64 relative gap of the two best codes < 1e-5)."""
import os

import numpy as np
import pytest
import torch

from hilcodec_b200 import _lib, checkpoint, fold, models
from hilcodec_b200 import streaming as S
from hilcodec_b200 import weights as W
from oracle import hilcodec_oracle as O

from helpers import GOLDEN, index_report, params, synth_wav

pytestmark = pytest.mark.gpu

TOL = 1e-4


def _fixture():
    return np.load(os.path.join(GOLDEN, "ref_train_random.npz"))


def _served(n_q, seed):
    """training-format checkpoint -> key rename -> fold (graph="train") -> CUDA model, plus the oracle's params."""
    cfg = W.CodecConfig(num_quantizers=n_q)
    sd = checkpoint.random_training_state_dict(cfg, seed)
    m = models.HILCodec.from_training_state_dict(sd, n_q).cuda()
    return m, params(checkpoint.deployment_weights(sd, cfg, "train")), O.CodecConfig(num_quantizers=n_q)


def test_reference_training_graph_fixture():
    g = _fixture()
    n_q, n = int(g["n_q"]), int(g["n"])
    m, p, ocfg = _served(n_q, int(g["seed"]))
    assert m.graph == "train" and m.deploy._core.graph == _lib.HIL_GRAPH_TRAIN
    for T in g["lengths"].tolist():
        x = torch.from_numpy(g[f"x_{T}"]).cuda()
        frames = -(-T // 320)
        z = m.encoder(x)
        assert z.shape == (x.shape[0], 128, frames)
        assert np.abs(z.cpu().numpy() - g[f"z_{T}"]).max() < TOL, f"latents, T={T}"
        q, num_replaces, loss, idx = m.quantizer(z, n, return_indices=True)
        assert idx.shape == (x.shape[0], n, frames) and idx.dtype == torch.int64
        assert num_replaces.shape == (n_q,) and not num_replaces.any()
        ref_idx = torch.from_numpy(g[f"indices_{T}"].astype(np.int64))
        bad, gap = index_report(ocfg, p, z.transpose(1, 2).contiguous(), idx.permute(1, 0, 2), ref_idx.permute(1, 0, 2), n)
        assert bad == 0 or gap < 1e-5, f"T={T}: {bad} frames disagree, worst relative gap {gap:.3e}"
        # decoder on the REFERENCE's quantized latents (independent of any near-tie above)
        y = m.decoder(torch.from_numpy(g[f"q_{T}"]).cuda())
        assert y.shape == (x.shape[0], 1, 320 * frames)
        assert np.abs(y.cpu().numpy() - g[f"wav_{T}"]).max() < TOL, f"decoder, T={T}"
        if bad == 0:
            wav, _, loss_vq = m(x, n)
            assert np.array_equal(q.cpu().numpy(), g[f"q_{T}"])  # gathers + in-order adds: bit-exact
            assert np.abs(wav.cpu().numpy() - g[f"wav_{T}"]).max() < TOL, f"forward, T={T}"
            assert abs(float(loss_vq) - float(g[f"loss_{T}"])) < 1e-4 and abs(float(loss) - float(g[f"loss_{T}"])) < 1e-4


@pytest.mark.parametrize("T", [1, 319, 321, 1000, 4001])
def test_ragged_encoder_against_oracle(T):
    """Every causal conv of the training graph pads its own right edge (modules/conv.py:61-68, :222-236)."""
    m, p, ocfg = _served(2, 11)
    x = synth_wav(3, T, seed=T)
    with torch.no_grad():
        ref = O.encoder_forward_train(ocfg, p, x)
    z = m.encoder(x.cuda())
    assert z.shape == ref.shape == (3, 128, -(-T // 320))
    assert (z.cpu() - ref).abs().max().item() < TOL


def test_ragged_call_equals_cached_call_on_hop_multiples():
    """With T a multiple of the hop nothing is padded: hil_encode_ragged must reproduce Encoder.forward with
    zero caches bit for bit, and leave the state reset."""
    w = W.random_weights(W.HIL_SPEECH, 13)
    m = S.HILCodec.from_weights(w, 8).cuda()
    x = synth_wav(2, 320 * 10, seed=5).cuda()
    z1, _ = m.encoder(x, *m.encoder.initialize_cache(x))
    z2 = m._core.encode_ragged(x)
    z3 = m._core.encode_ragged(x)
    assert torch.equal(z1, z2) and torch.equal(z2, z3)


def test_training_search_formula_against_oracle():
    """vector_quantize.py:146-152: argmin(-2 x.e + |e|^2), no |x|^2 term."""
    m, p, ocfg = _served(6, 17)
    g = torch.Generator().manual_seed(3)
    z = torch.nn.functional.normalize(torch.randn(4, 128, 150, generator=g), dim=1) * 128 ** 0.5
    with torch.no_grad():
        q_ref, loss_ref, idx_ref = O.rvq_forward_train(ocfg, p, z, 6)
    q, _, loss, idx = m.quantizer(z.cuda(), 6, return_indices=True)
    bad, gap = index_report(ocfg, p, z.transpose(1, 2).contiguous(), idx.permute(1, 0, 2), idx_ref.permute(1, 0, 2), 6)
    assert bad == 0 or gap < 1e-5, f"{bad} frames disagree, worst relative gap {gap:.3e}"
    if bad == 0:
        assert torch.equal(q.cpu(), q_ref)
        assert abs(float(loss) - float(loss_ref)) < 1e-5
    with pytest.raises(AssertionError):
        m.quantizer(z.cuda(), 7)


def test_graph_selection_contract():
    """The graph is part of the immutable model; the two graphs share weights but not the decoder arithmetic."""
    cfg = W.CodecConfig(num_quantizers=2)
    w = W.random_weights(cfg, 19)
    dep = S.HILCodec.from_weights(w, 2).cuda()
    tr = models.HILCodec(dep, graph="train").cuda()
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device())
    h_dep, h_tr = dep._core.model(dev), tr.deploy._core.model(dev)
    assert lib.hil_model_graph(h_dep) == _lib.HIL_GRAPH_DEPLOY and lib.hil_model_graph(h_tr) == _lib.HIL_GRAPH_TRAIN
    assert lib.hil_model_set_graph(h_dep, _lib.HIL_GRAPH_TRAIN) == -4  # HIL_ERR_STATE: finalized
    assert lib.hil_model_graph(h_dep) == _lib.HIL_GRAPH_DEPLOY
    q = torch.randn(2, 128, 6, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        ref_dep, _ = O.decoder_forward(O.CodecConfig(num_quantizers=2), params(w), q.transpose(1, 2))
        ref_tr, _ = O.decoder_forward(O.CodecConfig(num_quantizers=2), params(fold.to_train_graph(w)), q.transpose(1, 2),
                                      None, train_graph=True)
    y_dep = models.HILCodec(dep).decoder(q.cuda()).cpu()
    y_tr = tr.decoder(q.cuda()).cpu()
    assert (y_dep - ref_dep).abs().max().item() < TOL
    assert (y_tr - ref_tr).abs().max().item() < TOL
    assert (y_dep - y_tr).abs().max().item() > 1e-3  # quirks 1-2: not the same function
