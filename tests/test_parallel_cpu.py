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
    """Stand-in with the two gradient layouts of the sm_100a backward: flat-gradient mode (.grad are views of the
    persistent model.flat_grad) and autograd mode (.grad are separate tensors; flat_grad is a per-backward scratch)."""

    def __init__(self, rank, alias=True, accumulated=0.0):
        super().__init__()
        self.a = torch.nn.Parameter(torch.zeros(3, 4))
        self.b = torch.nn.Parameter(torch.zeros(5))
        self.flat_grad = torch.arange(20, dtype=torch.float32) * (rank + 1)   # 12 + pad to 15 + 5
        self.flat_grad_views = [self.flat_grad[0:12].view(3, 4), self.flat_grad[15:20]]
        # `accumulated`: what earlier micro-batches already left in .grad (autograd mode only)
        self.a.grad = self.flat_grad_views[0] if alias else self.flat_grad_views[0].clone() + accumulated
        self.b.grad = self.flat_grad_views[1] if alias else self.flat_grad_views[1].clone() + accumulated


def _grad_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ok = True
        mean = torch.arange(20, dtype=torch.float32) * (sum(range(1, world + 1)) / world)
        for alias, acc in ((True, 0.0), (False, 0.0), (False, 10.0)):
            m = _FlatGradModel(rank, alias, acc)
            aliased = allreduce_flat_gradients(m)
            # an accumulated .grad (two micro-batches) must be averaged as it is, never overwritten by the last step's buffer
            ok = ok and aliased == alias and torch.allclose(m.a.grad, mean[0:12].view(3, 4) + acc) and \
                torch.allclose(m.b.grad, mean[15:20] + acc)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_flat_gradient_allreduce_world2_gloo():
    """The training step's only collective: one all-reduce averages every parameter gradient on every rank — over the
    persistent flat buffer when the .grad tensors alias it, over the flattened .grad tensors (accumulation included)
    otherwise."""
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_grad_worker, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {0: True, 1: True}


def test_flat_gradient_allreduce_single_process():
    m = _FlatGradModel(0, alias=True)
    assert allreduce_flat_gradients(m) is True
    m2 = _FlatGradModel(0, alias=False, accumulated=3.0)
    before = m2.a.grad.clone()
    assert allreduce_flat_gradients(m2) is False and torch.equal(m2.a.grad, before)
