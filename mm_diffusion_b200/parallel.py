"""Batch-shard helpers for multi-GPU sampling and training (one process per GPU, torch.distributed).

The denoising path has no exchange inside a step: every rank holds a full weight replica and runs the
whole loop on its slice of the batch; the only collective is the gather of finished samples, which replaces the
per-rank file writes + barrier of the reference's sample script (py_scripts/multimodal_sample_sr.py:174-183,258) and
the all_gather in TrainLoop.save_video (mm_diffusion/multimodal_train_util.py:424-431)."""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.distributed as dist


def shard_bounds(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of the global batch owned by `rank` (sizes differ by at most one)."""
    base, extra = divmod(global_batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def rank_seed(base_seed: int, rank: int) -> int:
    """Independent RNG stream per rank (x_T, per-step noise, window shifts)."""
    return base_seed + 1000003 * rank


def to_uint8_video(video: torch.Tensor) -> torch.Tensor:
    """[-1,1] float video -> uint8, the conversion of multimodal_sample_sr.py:159-161."""
    return ((video + 1) * 127.5).clamp(0, 255).to(torch.uint8)


def sample_epilogue(video: torch.Tensor) -> torch.Tensor:
    """The whole video epilogue of the sampling scripts (multimodal_sample_sr.py:159-163) in one kernel: [B,F,C,H,W] float
    in [-1,1] -> uint8 [B,F,H,W,C] (`((v + 1) * 127.5).clamp(0, 255).to(uint8).permute(0, 1, 3, 4, 2).contiguous()`)."""
    if not video.is_cuda:
        return to_uint8_video(video).permute(0, 1, 3, 4, 2).contiguous()
    from . import _lib
    B, F, C, H, W = video.shape
    v = video.detach().to(torch.float32).contiguous()
    out = torch.empty((B, F, H, W, C), dtype=torch.uint8, device=video.device)
    with torch.cuda.device(video.device):
        _lib.check(_lib.load().mmd_sample_epilogue(v.data_ptr(), out.data_ptr(), B * F, C, H * W, _lib.current_stream_ptr()))
    return out


def gather_samples(sample: Dict[str, torch.Tensor], group=None) -> Dict[str, torch.Tensor]:
    """All ranks receive every rank's finished samples concatenated in rank order: uint8 video + fp32 audio
    (≈0.3 MB per sample).  Equal per-rank batch sizes are required (weak scaling: fixed batch per GPU)."""
    video = to_uint8_video(sample["video"]).contiguous()
    audio = sample["audio"].float().contiguous()
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return {"video": video, "audio": audio}
    world = dist.get_world_size(group)
    if dist.get_backend(group) != "nccl" and video.is_cuda:
        # gloo (CPU rendezvous, e.g. several ranks sharing one GPU in the tests): stage through the host
        lv = [torch.empty_like(video, device="cpu") for _ in range(world)]
        la = [torch.empty_like(audio, device="cpu") for _ in range(world)]
        dist.all_gather(lv, video.cpu(), group=group)
        dist.all_gather(la, audio.cpu(), group=group)
        return {"video": torch.cat(lv).to(video.device), "audio": torch.cat(la).to(audio.device)}
    gv = torch.empty((world * video.shape[0],) + tuple(video.shape[1:]), dtype=video.dtype, device=video.device)
    ga = torch.empty((world * audio.shape[0],) + tuple(audio.shape[1:]), dtype=audio.dtype, device=audio.device)
    dist.all_gather_into_tensor(gv, video, group=group)
    dist.all_gather_into_tensor(ga, audio, group=group)
    return {"video": gv, "audio": ga}


def _grads_alias(model) -> bool:
    """True when every existing .grad is the matching view of model.flat_grad (flat-gradient mode of the backward)."""
    flat = getattr(model, "flat_grad", None)
    views = getattr(model, "flat_grad_views", None)
    if flat is None or not views:
        return False
    for p, view in zip(model.parameters(), views):
        if p.grad is None:
            continue
        if view is None or p.grad.data_ptr() != view.data_ptr() or p.grad.shape != view.shape:
            return False
    return True


def allreduce_flat_gradients(model, group=None, average: bool = True) -> bool:
    """Data-parallel gradient exchange of a training step (reference: DistributedDataParallel in TrainLoop,
    mm_diffusion/multimodal_train_util.py:120-136).

    Flat-gradient mode (`model.use_flat_gradients(True)`): every `.grad` is a view of the model's persistent fp32
    buffer `model.flat_grad` (0.53 GB for the production network; micro-batches accumulate into it), so the exchange is
    ONE all-reduce over NCCL / NVLink on that buffer.  Returns True.

    Otherwise (gradients delivered through autograd: each `.grad` is its own tensor and may hold an accumulation over
    micro-batches) the `.grad` tensors themselves are flattened, reduced in one collective and written back — the
    accumulated values are what gets averaged, nothing is overwritten from a scratch buffer.  Returns False."""
    active = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    world = dist.get_world_size(group) if active else 1
    if _grads_alias(model):
        if active:
            dist.all_reduce(model.flat_grad, op=dist.ReduceOp.SUM, group=group)
            if average:
                model.flat_grad.div_(world)
        return True
    grads = [p.grad for p in model.parameters() if p.grad is not None]
    if not grads:
        raise RuntimeError("allreduce_flat_gradients: no parameter has a gradient (run backward first)")
    if active:
        from torch._utils import _flatten_dense_tensors, _unflatten_dense_tensors
        flat = _flatten_dense_tensors(grads)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            flat.div_(world)
        for g, r in zip(grads, _unflatten_dense_tensors(flat, grads)):
            g.copy_(r)
    return False
