"""BASELINE configs[4] "batch-sharded across 8 x B200 via NCCL", at world size 2 (-m gpu, needs two GPUs: run with
`gpurun --gpus 2`; skipped on a one-GPU box): every rank runs the CUDA path on its contiguous shard
(`sharding.forward_sharded`), the indices and PCM are all-gathered over NCCL back into batch order and must equal the
one-process result bit for bit (clips are independent; there is no collective inside the forward)."""
import os
import socket

import numpy as np
import pytest
import torch

from hilcodec_b200 import sharding
from hilcodec_b200 import weights as W

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist

    from hilcodec_b200 import streaming as S

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    cfg = W.HIL_MUSIC
    w = W.load_pretrained("hil_music") if W.have_pretrained("hil_music") else W.random_weights(cfg, 9)
    m = S.HILCodec.from_weights(w, 12).cuda()
    g = torch.Generator().manual_seed(3)
    x = (0.1 * torch.randn(5, 1, 320 * 25, generator=g)).clamp(-1, 1).to(dev)   # odd batch: shards of 3 and 2

    def compute(xs):
        return m.codec_forward(xs.contiguous(), 12)

    idx, wav = sharding.forward_sharded(compute, x, gather=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sharding.forward_sharded(lambda xs: (idx[:, :xs.shape[0]], wav[:xs.shape[0]]), x, gather=True)   # the collective alone
    e1.record()
    torch.cuda.synchronize()
    ms = sharding.reduce_max([e0.elapsed_time(e1)], device=dev)
    if rank == 0:
        full_idx, full_wav = compute(x)
        np.savez(os.path.join(out_dir, "r.npz"), same_idx=bool(torch.equal(idx, full_idx)),
                 same_wav=bool(torch.equal(wav, full_wav)), shape=np.array(idx.shape), gather_ms=ms[0])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_two_rank_nccl_gather_matches_single_process(tmp_path):
    import torch.multiprocessing as mp

    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    r = np.load(os.path.join(str(tmp_path), "r.npz"))
    assert list(r["shape"]) == [12, 5, 25]
    assert bool(r["same_idx"]) and bool(r["same_wav"])
    print("NCCL all-gather of idx + wav (2 ranks, 5 clips):", float(r["gather_ms"]), "ms")
