"""Sampler / loss parity on the GPU through the public API (create_gaussian_diffusion + MultimodalUNet shim) against
fixtures from the unmodified reference (tests/golden/small.pt) and against the CPU oracle.
Tolerances (fp16 network vs fp32 reference): rel-L2 <= 2e-2 on eps-derived quantities, <= 2e-3 on x_{t-1} for t >= 100
(the sample is dominated by x_t and the injected noise), loss rel-err <= 1e-2."""
import random

import pytest
import torch

from oracle.mmdiff_oracle import DiffusionOracle, draw_shifts, synthetic_state_dict
from tests.util_golden import build_b200_model, cfg_of, golden_inputs, load_golden, rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup():
    from mm_diffusion_b200.script_util import create_gaussian_diffusion
    fx = load_golden("small")
    cfg = cfg_of(fx)
    sd = synthetic_state_dict(cfg, seed=fx["weight_seed"])
    model = build_b200_model(cfg, sd)
    return fx, cfg, sd, model, create_gaussian_diffusion()


def _noise(fx, v, a):
    g = torch.Generator().manual_seed(fx["noise_seed"])
    return torch.randn(v.shape, generator=g), torch.randn(a.shape, generator=g)


def test_p_sample_matches_reference_golden(setup):
    fx, cfg, sd, model, diffusion = setup
    v, a = golden_inputs(cfg, fx)
    zv, za = _noise(fx, v, a)
    for tag in ("mid", "zero"):
        g_ = fx[f"p_sample_{tag}"]
        random.seed(11)
        with torch.no_grad():
            out = diffusion.p_sample(model, {"video": v.cuda(), "audio": a.cuda()}, g_["t"].cuda(),
                                     noise={"video": zv.cuda(), "audio": za.cuda()})
        errs = {k: rel_l2(out[grp][mod], g_[f"{key}_{mod}"]) for grp, key in (("sample", "sample"), ("pred_start", "pred_start"))
                for mod in ("video", "audio") for k in [f"{key}_{mod}"]}
        print(tag, {k: f"{e:.2e}" for k, e in errs.items()})
        assert all(e < 2e-2 for e in errs.values())
        assert errs["sample_video"] < 5e-3 and errs["sample_audio"] < 5e-3


def test_p_sample_draws_noise_like_reference(setup):
    """Without injected noise the step draws th.randn_like(video) then th.randn_like(audio) from the CUDA generator."""
    fx, cfg, sd, model, diffusion = setup
    v, a = golden_inputs(cfg, fx)
    x = {"video": v.cuda(), "audio": a.cuda()}
    t = torch.tensor([400, 400]).cuda()
    shifts = draw_shifts(cfg, random.Random(2))
    with torch.no_grad():
        torch.manual_seed(77)
        zv = torch.randn_like(x["video"]); za = torch.randn_like(x["audio"])
        random.seed(21)
        ref = diffusion.p_sample(model, x, t, noise={"video": zv, "audio": za})
        torch.manual_seed(77)
        random.seed(21)
        out = diffusion.p_sample(model, x, t)
    assert rel_l2(out["sample"]["video"], ref["sample"]["video"]) < 5e-3
    assert rel_l2(out["sample"]["audio"], ref["sample"]["audio"]) < 5e-3


def test_short_loop_matches_oracle(setup):
    """5 chained ancestral steps (t = 999..995) with injected noise vs the CPU oracle; the window shifts come from the
    same seeded global `random` stream on both sides (the model draws them like the reference does)."""
    fx, cfg, sd, model, diffusion = setup
    oracle = DiffusionOracle(1000)
    g = torch.Generator().manual_seed(314)
    B = 2
    x = {"video": torch.randn(B, *cfg.video_size, generator=g), "audio": torch.randn(B, *cfg.audio_size, generator=g)}
    xg = {k: val.cuda() for k, val in x.items()}
    for i in range(999, 994, -1):
        z = {"video": torch.randn(B, *cfg.video_size, generator=g), "audio": torch.randn(B, *cfg.audio_size, generator=g)}
        t = torch.full((B,), i, dtype=torch.long)
        with torch.no_grad():
            random.seed(1000 + i)
            x = oracle.p_sample(sd, cfg, x, t, z, draw_shifts(cfg, random))["sample"]
            random.seed(1000 + i)
            xg = diffusion.p_sample(model, xg, t.cuda(), noise={k: val.cuda() for k, val in z.items()})["sample"]
    ev, ea = rel_l2(xg["video"], x["video"]), rel_l2(xg["audio"], x["audio"])
    print(f"5-step loop rel-L2 video {ev:.2e} audio {ea:.2e}")
    assert ev < 5e-3 and ea < 5e-3


def test_q_sample_and_training_losses(setup):
    fx, cfg, sd, model, diffusion = setup
    v, a = golden_inputs(cfg, fx)
    zv, za = _noise(fx, v, a)
    tr = fx["training"]
    oracle = DiffusionOracle(1000)
    qs = diffusion.q_sample(v.clamp(-1, 1).cuda(), tr["t"].cuda(), noise=zv.cuda())
    assert rel_l2(qs, oracle.q_sample(v.clamp(-1, 1), tr["t"], zv)) < 1e-6
    random.seed(13)
    with torch.no_grad():
        losses = diffusion.multimodal_training_losses(model, {"video": v.clamp(-1, 1).cuda(), "audio": a.clamp(-1, 1).cuda()},
                                                      tr["t"].cuda(), noise={"video": zv.cuda(), "audio": za.cuda()})
    for k in ("loss", "mse_video", "mse_audio"):
        err = ((losses[k].cpu() - tr[k]).abs() / tr[k].abs()).max().item()
        print(k, f"rel err {err:.2e}")
        assert err < 1e-2


def test_p_sample_loop_api_and_replacement_conditioning(setup):
    """p_sample_loop over a respaced 4-step schedule returns the reference's dict layout; the replacement-method
    conditional loop keeps the conditioned modality on the q(x_t | condition) trajectory."""
    from mm_diffusion_b200.script_util import create_gaussian_diffusion
    fx, cfg, sd, model, _ = setup
    diff = create_gaussian_diffusion(timestep_respacing="4")
    B = 2
    shape = {"video": (B, *cfg.video_size), "audio": (B, *cfg.audio_size)}
    torch.manual_seed(0)
    out = diff.p_sample_loop(model, shape, clip_denoised=True, device=torch.device("cuda"), progress=False)
    assert set(out) == {"video", "audio"} and out["video"].shape == shape["video"] and torch.isfinite(out["video"]).all()
    cond = torch.rand(B, *cfg.audio_size) * 0.2 - 0.1
    torch.manual_seed(1)
    out = diff.conditional_p_sample_loop(model, shape, use_fp16=True, model_kwargs={"audio": cond.cuda()},
                                         device=torch.device("cuda"), progress=False, class_scale=0.0)
    assert out["video"].shape == shape["video"] and torch.isfinite(out["audio"]).all()
    # gradient guidance (class_scale > 0, reference :722-819) backpropagates through the sm_100a path to the video input
    torch.manual_seed(2)
    out = diff.conditional_p_sample_loop(model, shape, use_fp16=False, model_kwargs={"audio": cond.cuda()},
                                         device=torch.device("cuda"), progress=False, class_scale=3.0)
    assert out["video"].shape == shape["video"] and torch.isfinite(out["video"]).all() and torch.isfinite(out["audio"]).all()


def test_sample_epilogue_matches_the_script_conversion():
    """uint8 + channels-last epilogue of multimodal_sample_sr.py:159-163 in one kernel, byte for byte."""
    from mm_diffusion_b200.parallel import sample_epilogue
    g = torch.Generator().manual_seed(5)
    v = (torch.randn(3, 16, 3, 64, 64, generator=g) * 0.8).cuda()
    v[0, 0, 0, 0, :4] = torch.tensor([-1.0, 1.0, -3.0, 3.0]).cuda()
    ref = ((v + 1) * 127.5).clamp(0, 255).to(torch.uint8).permute(0, 1, 3, 4, 2).contiguous()
    got = sample_epilogue(v)
    assert got.dtype == torch.uint8 and got.shape == ref.shape and torch.equal(got, ref)
