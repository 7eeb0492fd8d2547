"""CPU: the product's zero-shot conditional sampling loops (mm_diffusion_b200/gaussian_diffusion.py, host logic of
SURVEY.md §8 row a19) against fixtures from the UNMODIFIED reference loops (oracle/make_golden_cond.py ->
tests/golden/cond_small.pt; multimodal_gaussian_diffusion.py:584-819).  The model is the CPU oracle (a plain
differentiable torch callable), so this pins the loop arithmetic, the order of the RNG draws (x_T video then audio, per-step
noise video then audio, one randint per shifting block) and the guidance update, independent of any kernel."""
import random

import pytest
import torch

from mm_diffusion_b200.script_util import create_gaussian_diffusion
from oracle.mmdiff_oracle import draw_shifts, synthetic_state_dict, unet_forward
from tests.util_golden import cfg_of, load_golden, rel_l2

FX = load_golden("cond_small")


class _OracleModel:
    """Reference-surface callable over the CPU oracle: draws its window shifts from the global `random` like the model."""

    def __init__(self, cfg, sd):
        self.cfg, self.sd = cfg, sd

    def __call__(self, video, audio, t, **kw):
        return unet_forward(self.sd, self.cfg, video, audio, t, draw_shifts(self.cfg, random))


@pytest.mark.parametrize("name", sorted(FX["cases"]))
def test_conditional_loops_match_reference(name):
    case = FX["cases"][name]
    cfg = cfg_of(FX)
    sd = synthetic_state_dict(cfg, seed=FX["weight_seed"])
    B = FX["batch"]
    shape = {"video": (B, *cfg.video_size), "audio": (B, *cfg.audio_size)}
    g = torch.Generator().manual_seed(FX["cond_seed"])
    cond = {"video": 0.1 * torch.randn(B, *cfg.video_size, generator=g), "audio": 0.1 * torch.randn(B, *cfg.audio_size, generator=g)}
    diffusion = create_gaussian_diffusion(timestep_respacing=FX["respacing"])
    torch.manual_seed(case["torch_seed"])
    random.seed(case["shift_seed"])
    orig = random.randint
    if case["const_shift"] is not None:   # see oracle/make_golden_cond.py: the reference re-draws shifts in its backward
        random.randint = lambda lo, hi: min(hi, case["const_shift"])
    try:
        out = diffusion.conditional_p_sample_loop(_OracleModel(cfg, sd), shape, case["use_fp16"], clip_denoised=True,
                                                  model_kwargs={case["condition"]: cond[case["condition"]].clone()},
                                                  progress=False, class_scale=case["class_scale"], device=torch.device("cpu"))
    finally:
        random.randint = orig
    ev, ea = rel_l2(out["video"], case["video"]), rel_l2(out["audio"], case["audio"])
    print(name, f"video {ev:.2e} audio {ea:.2e}")
    assert ev < 1e-4 and ea < 1e-4
