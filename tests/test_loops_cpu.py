"""CPU: the product's sampling loops (p_sample_loop, ddim_sample_loop; mm_diffusion_b200/gaussian_diffusion.py and
respace.py, host logic of SURVEY.md §8 rows a16-a18, a22) against fixtures from the UNMODIFIED reference loops
(oracle/make_golden_loops.py -> tests/golden/loops_small.pt).  The model is the CPU oracle as a plain callable, so
this pins schedule respacing, timestep mapping, the order of the RNG draws and the update arithmetic."""
import random

import pytest
import torch

from mm_diffusion_b200.script_util import create_gaussian_diffusion
from oracle.mmdiff_oracle import draw_shifts, synthetic_state_dict, unet_forward
from tests.util_golden import cfg_of, load_golden, rel_l2

FX = load_golden("loops_small")


@pytest.mark.parametrize("name", sorted(FX["cases"]))
def test_sampling_loops_match_reference(name):
    case = FX["cases"][name]
    cfg = cfg_of(FX)
    sd = synthetic_state_dict(cfg, seed=FX["weight_seed"])
    B = FX["batch"]
    shape = {"video": (B, *cfg.video_size), "audio": (B, *cfg.audio_size)}
    diffusion = create_gaussian_diffusion(timestep_respacing=case["respacing"])
    assert list(diffusion.timestep_map) == case["timestep_map"]

    def model(video, audio, t, **kw):
        return unet_forward(sd, cfg, video, audio, t, draw_shifts(cfg, random))

    torch.manual_seed(case["torch_seed"])
    random.seed(case["shift_seed"])
    with torch.no_grad():
        out = getattr(diffusion, case["loop"])(model, shape, progress=False, device=torch.device("cpu"), **case["kwargs"])
    ev, ea = rel_l2(out["video"], case["video"]), rel_l2(out["audio"], case["audio"])
    print(name, f"video {ev:.2e} audio {ea:.2e}")
    assert ev < 1e-4 and ea < 1e-4
