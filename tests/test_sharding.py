"""N > 1 host logic on CPU: two gloo processes shard a batch, run the codec path on their shard
(the CPU oracle stands in for the CUDA kernels here) and must reproduce the single-process
result in batch order; timing counters reduce with MAX."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hilcodec_b200 import sharding
from hilcodec_b200 import weights as W


def test_shard_bounds_cover_batch_exactly():
    for total in (0, 1, 7, 256, 2048):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert sharding.shard_bounds(2048, 8, 3) == (768, 1024)  # BASELINE config 5: 256 clips per GPU
    with pytest.raises(ValueError):
        sharding.shard_bounds(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from oracle import hilcodec_oracle as O

    cfg = W.CodecConfig(num_quantizers=4)
    p = {k: torch.from_numpy(v) for k, v in W.random_weights(cfg, 9).items()}
    ocfg = O.CodecConfig(num_quantizers=4)
    g = torch.Generator().manual_seed(3)
    x = (0.1 * torch.randn(3, 1, 320 * 4, generator=g)).clamp(-1, 1)  # odd batch: shards of 2 and 1

    def compute(xs):
        with torch.no_grad():
            o = O.codec_forward(ocfg, p, xs, 4)
        return o["indices"], o["wav"]

    idx, wav = sharding.forward_sharded(compute, x, gather=True)
    ms = sharding.reduce_max([10.0 + rank, 5.0 - rank])
    tot = sharding.reduce_sum([float(sharding.shard_bounds(3, world, rank)[1] - sharding.shard_bounds(3, world, rank)[0])])
    if rank == 0:
        full_idx, full_wav = compute(x)
        np.savez(os.path.join(out_dir, "r.npz"), same_idx=bool(torch.equal(idx, full_idx)),
                 wav_err=float((wav - full_wav).abs().max()), ms=np.array(ms), tot=tot[0])
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_process(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r = np.load(os.path.join(str(tmp_path), "r.npz"))
    assert bool(r["same_idx"])
    assert float(r["wav_err"]) < 1e-6
    assert list(r["ms"]) == [11.0, 5.0]
    assert float(r["tot"]) == 3.0
