"""world_size-2 gloo test (CPU) of the N>1 host logic: batch sharding, per-rank seeds, and the single gather of
finished samples that the batch-sharded sampling loop ends with."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mm_diffusion_b200.parallel import gather_samples, rank_seed, shard_bounds, to_uint8_video


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        per = 2
        g = torch.Generator().manual_seed(rank_seed(1234, rank))
        sample = {"video": torch.randn(per, 4, 3, 8, 8, generator=g).clamp(-1, 1), "audio": torch.randn(per, 1, 64, generator=g)}
        out = gather_samples(sample)
        # every rank must hold rank-ordered concatenation
        expect_v, expect_a = [], []
        for r in range(world):
            gg = torch.Generator().manual_seed(rank_seed(1234, r))
            expect_v.append(to_uint8_video(torch.randn(per, 4, 3, 8, 8, generator=gg).clamp(-1, 1)))
            expect_a.append(torch.randn(per, 1, 64, generator=gg))
        ok = torch.equal(out["video"], torch.cat(expect_v)) and torch.equal(out["audio"], torch.cat(expect_a))
        ret[rank] = bool(ok) and out["video"].dtype == torch.uint8
    finally:
        dist.destroy_process_group()


def test_gather_samples_world2_gloo():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {0: True, 1: True}


def test_shard_bounds_partition_the_batch():
    for gb in (1, 4, 7, 32):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds(gb, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == gb
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert len({rank_seed(7, r) for r in range(8)}) == 8
