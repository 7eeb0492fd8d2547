"""CPU: pin the oracle (oracle/mmdiff_oracle.py) against fixtures produced by the unmodified reference
(oracle/make_golden.py -> tests/golden/*.pt).  fp32 vs fp32 -> tight tolerances (summation order only)."""
import os
import random

import pytest
import torch

from oracle.mmdiff_oracle import (DiffusionOracle, UNetConfig, draw_shifts, param_shapes, shift_bounds,
                                  synthetic_state_dict, unet_forward)

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    path = os.path.join(GOLDEN, name + ".pt")
    if not os.path.exists(path):
        pytest.skip(f"{path} not generated")
    return torch.load(path, weights_only=False)


def cfg_of(fx):
    return UNetConfig(**fx["config"])


def inputs(cfg, fx):
    g = torch.Generator().manual_seed(fx["input_seed"])
    v = torch.randn(fx["batch"], *cfg.video_size, generator=g)
    a = torch.randn(fx["batch"], *cfg.audio_size, generator=g)
    return v, a


def rel(a, b):
    return ((a - b).norm() / b.norm()).item()


@pytest.mark.parametrize("name", ["small", "production"])
def test_param_inventory_matches_reference(name):
    fx = load(name)
    cfg = cfg_of(fx)
    mine = param_shapes(cfg)
    ref = [(k, tuple(s)) for k, s in fx["state_dict_keys"]]
    assert mine == ref  # same names, shapes and registration order
    assert sum(torch.Size(s).numel() for _, s in mine) == fx["num_params"]


@pytest.mark.parametrize("name", ["small", "production"])
def test_shift_draws_match_reference(name):
    fx = load(name)
    cfg = cfg_of(fx)
    draws = fx["forward_shift_draws"]
    bounds = [b for b in shift_bounds(cfg) if b >= 0]
    assert [(lo, hi) for lo, hi, _ in draws] == [(0, b) for b in bounds]
    random.seed(7)
    mine = [s for s, b in zip(draw_shifts(cfg, random), shift_bounds(cfg)) if b >= 0]
    assert mine == [v for _, _, v in draws]


def test_forward_small_matches_reference():
    fx = load("small")
    cfg = cfg_of(fx)
    sd = synthetic_state_dict(cfg, seed=fx["weight_seed"])
    v, a = inputs(cfg, fx)
    random.seed(7)
    shifts = draw_shifts(cfg, random)
    with torch.no_grad():
        ev, ea = unet_forward(sd, cfg, v, a, fx["t"], shifts)
    assert rel(ev, fx["forward_video"]) < 2e-5
    assert rel(ea, fx["forward_audio"]) < 2e-5


def test_p_sample_and_loss_small_match_reference():
    fx = load("small")
    cfg = cfg_of(fx)
    sd = synthetic_state_dict(cfg, seed=fx["weight_seed"])
    v, a = inputs(cfg, fx)
    g = torch.Generator().manual_seed(fx["noise_seed"])
    zv = torch.randn(v.shape, generator=g)
    za = torch.randn(a.shape, generator=g)
    diff = DiffusionOracle(1000)
    assert torch.allclose(torch.from_numpy(diff.alphas_cumprod), fx["alphas_cumprod"], rtol=0, atol=0)
    assert torch.allclose(torch.from_numpy(diff.posterior_mean_coef1), fx["posterior_mean_coef1"], rtol=0, atol=0)
    for tag in ("mid", "zero"):
        g_ = fx[f"p_sample_{tag}"]
        random.seed(11)
        shifts = draw_shifts(cfg, random)
        with torch.no_grad():
            out = diff.p_sample(sd, cfg, {"video": v, "audio": a}, g_["t"], {"video": zv, "audio": za}, shifts)
        assert rel(out["sample"]["video"], g_["sample_video"]) < 2e-5
        assert rel(out["sample"]["audio"], g_["sample_audio"]) < 2e-5
        assert rel(out["pred_start"]["video"], g_["pred_start_video"]) < 2e-5
        assert rel(out["pred_start"]["audio"], g_["pred_start_audio"]) < 2e-5
    tr = fx["training"]
    random.seed(13)
    shifts = draw_shifts(cfg, random)
    with torch.no_grad():
        losses = diff.training_losses(sd, cfg, {"video": v.clamp(-1, 1), "audio": a.clamp(-1, 1)}, tr["t"],
                                      {"video": zv, "audio": za}, shifts)
    for k in ("loss", "mse_video", "mse_audio"):
        assert rel(losses[k], tr[k]) < 2e-5


@pytest.mark.slow
def test_forward_production_matches_reference():
    fx = load("production")
    cfg = cfg_of(fx)
    sd = synthetic_state_dict(cfg, seed=fx["weight_seed"])
    v, a = inputs(cfg, fx)
    random.seed(7)
    shifts = draw_shifts(cfg, random)
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        ev, ea = unet_forward(sd, cfg, v, a, fx["t"], shifts)
    assert rel(ev.flatten()[::8], fx["forward_video_sub"]) < 5e-5
    assert rel(ea.flatten()[::8], fx["forward_audio_sub"]) < 5e-5
    assert abs(ev.norm().item() - fx["forward_video_norm"]) / fx["forward_video_norm"] < 5e-5
