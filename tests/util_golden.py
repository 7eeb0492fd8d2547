"""Shared helpers for the parity tests: golden fixtures (generated from the unmodified reference by
oracle/make_golden.py) and model construction with the oracle's deterministic synthetic weights."""
import os

import torch

from oracle.mmdiff_oracle import UNetConfig, synthetic_state_dict

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)


def cfg_of(fx):
    return UNetConfig(**fx["config"])


def golden_inputs(cfg, fx):
    g = torch.Generator().manual_seed(fx["input_seed"])
    v = torch.randn(fx["batch"], *cfg.video_size, generator=g)
    a = torch.randn(fx["batch"], *cfg.audio_size, generator=g)
    return v, a


def build_b200_model(cfg: UNetConfig, sd=None, device="cuda", **kw):
    from mm_diffusion_b200.unet import MultimodalUNet
    m = MultimodalUNet(list(cfg.video_size), list(cfg.audio_size), cfg.model_channels, cfg.video_out_channels,
                       cfg.audio_out_channels, cfg.num_res_blocks, list(cfg.cross_attention_resolutions),
                       list(cfg.cross_attention_windows), cfg.cross_attention_shift,
                       list(cfg.video_attention_resolutions), list(cfg.audio_attention_resolutions),
                       channel_mult=tuple(cfg.channel_mult), num_heads=cfg.num_heads,
                       num_head_channels=cfg.num_head_channels, use_scale_shift_norm=True, resblock_updown=True, **kw)
    if sd is not None:
        m.load_state_dict(sd, strict=True)
    return m.to(device).eval()


def rel_l2(a, b):
    a = a.detach().float().cpu()
    b = b.detach().float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()
