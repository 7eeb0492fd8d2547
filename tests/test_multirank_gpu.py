"""Multi-rank sampling on the GPU: N ranks, each running the whole loop on its own batch shard
with its own RNG stream, must produce — after the one gather of finished samples — exactly the bytes that N
single-process runs with those seeds produce.  One process per rank; with fewer GPUs than ranks the ranks share cuda:0
and rendezvous over gloo (NCCL refuses two ranks on one device), otherwise one GPU per rank over NCCL."""
import os
import random
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

WORLD = 2
PER_RANK = 2
STEPS = "4"


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _sample(model, diffusion, cfg, seed, device):
    """One short respaced ancestral loop with every random source pinned to `seed` (x_T and per-step noise from
    torch's generators, window shifts from Python's `random`, like the reference's loop)."""
    from mm_diffusion_b200.parallel import rank_seed  # noqa: F401
    torch.manual_seed(seed)
    torch.cuda.manual_seed(seed)
    random.seed(seed)
    shape = {"video": (PER_RANK, *cfg.video_size), "audio": (PER_RANK, *cfg.audio_size)}
    with torch.no_grad():
        return diffusion.p_sample_loop(model, shape, clip_denoised=True, device=device, progress=False)


def _build(device):
    from mm_diffusion_b200.script_util import create_gaussian_diffusion
    from oracle.make_golden import SMALL as cfg
    from oracle.mmdiff_oracle import synthetic_state_dict
    from tests.util_golden import build_b200_model
    model = build_b200_model(cfg, synthetic_state_dict(cfg, seed=0), device=device)
    return cfg, model, create_gaussian_diffusion(steps=1000, timestep_respacing=STEPS)


def _worker(rank, world, port, backend, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = torch.device("cuda", rank if backend == "nccl" else 0)
    torch.cuda.set_device(dev)
    dist.init_process_group(backend, rank=rank, world_size=world)
    try:
        from mm_diffusion_b200.parallel import gather_samples, rank_seed
        cfg, model, diffusion = _build(dev)
        mine = _sample(model, diffusion, cfg, rank_seed(77, rank), dev)
        got = gather_samples(mine)
        if rank == 0:
            # the same process replays every rank's seed alone and compares bytes
            ok = True
            for r in range(world):
                ref = gather_samples_single(_sample(model, diffusion, cfg, rank_seed(77, r), dev))
                sl = slice(r * PER_RANK, (r + 1) * PER_RANK)
                ok = ok and torch.equal(got["video"][sl], ref["video"]) and torch.equal(got["audio"][sl], ref["audio"])
            ret["ok"] = bool(ok)
            ret["shape"] = tuple(got["video"].shape)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def gather_samples_single(sample):
    from mm_diffusion_b200.parallel import to_uint8_video
    return {"video": to_uint8_video(sample["video"]).contiguous(), "audio": sample["audio"].float().contiguous()}


def test_ranks_reproduce_single_process_runs_byte_for_byte():
    backend = "nccl" if torch.cuda.device_count() >= WORLD else "gloo"
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(WORLD, port, backend, ret), nprocs=WORLD, join=True)
        assert ret.get("ok") is True, dict(ret)
        assert ret["shape"][0] == WORLD * PER_RANK
