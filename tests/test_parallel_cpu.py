"""world_size-2 gloo test (CPU) of the N>1 host logic: batch sharding, per-rank seeds, and the single gather of
finished samples that the batch-sharded sampling loop ends with."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mm_diffusion_b200.parallel import allreduce_flat_gradients, gather_samples, rank_seed, shard_bounds, to_uint8_video


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


class _FlatGradModel(torch.nn.Module):
    """Stand-in with the gradient layout the sm_100a backward produces: .grad tensors are views of one flat buffer."""

    def __init__(self, rank, alias=True):
        super().__init__()
        self.a = torch.nn.Parameter(torch.zeros(3, 4))
        self.b = torch.nn.Parameter(torch.zeros(5))
        self.flat_grad = torch.arange(20, dtype=torch.float32) * (rank + 1)   # 12 + pad to 15 + 5
        self.flat_grad_views = [self.flat_grad[0:12].view(3, 4), self.flat_grad[15:20]]
        self.a.grad = self.flat_grad_views[0] if alias else self.flat_grad_views[0].clone()
        self.b.grad = self.flat_grad_views[1] if alias else self.flat_grad_views[1].clone()


def _grad_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ok = True
        for alias in (True, False):
            m = _FlatGradModel(rank, alias)
            aliased = allreduce_flat_gradients(m)
            mean = torch.arange(20, dtype=torch.float32) * (sum(range(1, world + 1)) / world)
            ok = ok and aliased == alias and torch.allclose(m.a.grad, mean[0:12].view(3, 4)) and torch.allclose(m.b.grad, mean[15:20])
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_flat_gradient_allreduce_world2_gloo():
    """The training step's only collective: one all-reduce over the flat gradient buffer averages every parameter
    gradient on every rank (views alias the buffer; non-aliased .grad tensors are refreshed by copy)."""
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_grad_worker, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {0: True, 1: True}
