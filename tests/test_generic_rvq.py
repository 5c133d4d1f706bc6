"""SURVEY.md section 8a rows Q'' / V'': the GENERIC quantizer `modules/vector_quantize.py:471 ResidualVQ` (the class
`north_star` names), `VectorQuantize` :376 and `EuclideanCodebook` :76 -- drop-in `hilcodec_b200.vector_quantize`.
Pinned to `tests/golden/ref_generic_rvq.npz`, made by the reference's own class (`make_golden.py generic_rvq`)."""
import os
import sys

import numpy as np
import pytest
import torch

from hilcodec_b200 import vector_quantize as VQ
from oracle import hilcodec_oracle as O
from oracle import ref_shim

from helpers import GOLDEN

sys.path.insert(0, GOLDEN)
from make_golden import generic_rvq_inputs  # noqa: E402

need_ref = pytest.mark.skipif(not ref_shim.available(), reason="reference tree absent (GPU box)")


def _fixture():
    g = np.load(os.path.join(GOLDEN, "ref_generic_rvq.npz"))
    embeds, x = generic_rvq_inputs(int(g["seed"]), int(g["n_q"]), int(g["size"]), int(g["batch"]), int(g["frames"]))
    return g, embeds, x


CASES = [(cl, n) for cl in (False, True) for n in (None, 3, 1)]


@pytest.mark.parametrize("channel_last,n", CASES)
def test_oracle_matches_reference_fixture(channel_last, n):
    g, embeds, x = _fixture()
    xin = torch.from_numpy(x if not channel_last else np.ascontiguousarray(x.transpose(0, 2, 1)))
    q, loss, idx = O.generic_rvq_forward([torch.from_numpy(e) for e in embeds], xin, n, channel_last)
    tag = f"{'cl' if channel_last else 'cf'}_{n}"
    assert np.array_equal(q.numpy(), g[f"q_{tag}"])
    assert abs(float(loss) - float(g[f"loss_{tag}"])) < 1e-6
    assert idx.shape == (int(g["n_q"]) if n is None else n, x.shape[0], x.shape[2])


@need_ref
def test_state_dict_keys_and_constructor_match_reference():
    ref_shim.import_streaming()
    from modules.vector_quantize import ResidualVQ as RefRVQ  # type: ignore
    kw = dict(dim=128, codebook_size=64, kmeans_init=False, decay=0.9, ema_num_threshold=0.5, channel_last=True,
              commitment=0.25)
    ref = RefRVQ(4, dropout=True, dropout_index=[2, 4], **kw)
    mine = VQ.ResidualVQ(4, dropout=True, dropout_index=[2, 4], **kw)
    rs, ms = ref.state_dict(), mine.state_dict()
    assert list(rs) == list(ms)
    assert all(rs[k].shape == ms[k].shape and rs[k].dtype == ms[k].dtype for k in rs)
    mine.load_state_dict(rs)   # strict
    assert all(torch.equal(mine.state_dict()[k], rs[k]) for k in rs)
    assert mine.dropout_index == ref.dropout_index and mine.use_shape_gain == ref.use_shape_gain
    assert mine.layers[0].commitment == ref.layers[0].commitment and mine.layers[0].channel_last is True


def test_scope_errors_without_gpu():
    vq = VQ.ResidualVQ(2, dim=128, codebook_size=16)
    x = torch.zeros(1, 128, 4)
    with pytest.raises(NotImplementedError):      # constructed in training mode, like any nn.Module
        vq(x)
    vq.eval()
    with pytest.raises(AssertionError):           # modules/vector_quantize.py:496-497
        vq(x, 3)
    with pytest.raises(RuntimeError):             # no CPU path
        vq(x)
    with pytest.raises(NotImplementedError):
        VQ.ResidualVQ(2, dim=128, codebook_size=16, use_shape_gain=True)
    with pytest.raises(NotImplementedError):
        VQ.ResidualVQ(2, dim=64, codebook_size=16)
    with pytest.raises(NotImplementedError):      # k-means initialisation is a training-time step
        VQ.EuclideanCodebook(128, 16, kmeans_init=True).eval()(torch.zeros(1, 2, 128))


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("channel_last,n", CASES)
def test_gpu_residual_vq_matches_reference_fixture(channel_last, n):
    g, embeds, x = _fixture()
    vq = VQ.ResidualVQ(int(g["n_q"]), dim=128, codebook_size=int(g["size"]), channel_last=channel_last).eval()
    sd = vq.state_dict()
    for i, e in enumerate(embeds):
        sd[f"layers.{i}._codebook.embed"] = torch.from_numpy(e)
    vq.load_state_dict(sd)
    vq = vq.cuda()
    xin = torch.from_numpy(x if not channel_last else np.ascontiguousarray(x.transpose(0, 2, 1))).cuda()
    q, num_replaces, loss = vq(xin, n)
    tag = f"{'cl' if channel_last else 'cf'}_{n}"
    assert q.shape == xin.shape and q.dtype == torch.float32
    assert num_replaces.dtype == np.int64 and num_replaces.shape == (int(g["n_q"]),) and not num_replaces.any()
    assert np.array_equal(q.cpu().numpy(), g[f"q_{tag}"])          # gathers + in-order fp32 adds: bit-exact
    assert abs(float(loss) - float(g[f"loss_{tag}"])) < 1e-6


@pytest.mark.gpu
def test_gpu_vector_quantize_and_codebook_single_stage():
    g, embeds, x = _fixture()
    e0 = torch.from_numpy(embeds[0])
    xt = torch.from_numpy(np.ascontiguousarray(x.transpose(0, 2, 1)))
    q_ref, idx_ref = O.codebook_search(xt, e0)
    cb = VQ.EuclideanCodebook(128, e0.shape[0]).eval()
    cb.embed.copy_(e0)
    cb = cb.cuda()
    q, num_replace = cb(xt.cuda())
    assert num_replace == 0 and torch.equal(q.cpu(), q_ref)
    q4, _ = cb(xt.cuda().view(3, 5, 10, 128))                      # '... d' inputs keep their shape (:144, :159)
    assert q4.shape == (3, 5, 10, 128) and torch.equal(q4.cpu().view_as(q_ref), q_ref)
    layer = VQ.VectorQuantize(dim=128, codebook_size=e0.shape[0], commitment=0.5).eval()
    layer._codebook.embed.copy_(e0)
    layer = layer.cuda()
    xc = torch.from_numpy(x).cuda()
    ql, nr, commit = layer(xc, calculate_commitment_loss=True)
    assert nr == 0 and torch.equal(ql.cpu(), q_ref.transpose(1, 2))
    want = torch.nn.functional.mse_loss(q_ref, xt) * 0.5
    assert abs(float(commit) - float(want)) < 1e-6
    assert layer(xc)[2] is None
    # a codebook written in place after the first call is picked up (the native copy follows the buffer)
    with torch.no_grad():
        layer._codebook.embed.mul_(0.5)
    q2, _, _ = layer(xc)
    q2_ref, _ = O.codebook_search(xt, e0 * 0.5)
    assert torch.equal(q2.cpu(), q2_ref.transpose(1, 2))
