"""Generate tests/golden/loops_small.pt by running the UNMODIFIED reference sampling loops
(/root/reference/mm_diffusion/multimodal_gaussian_diffusion.py: p_sample_loop :476-582, ddim_sample_loop :955-1060)
on the SMALL reference model, CPU fp32, with respaced schedules.

TEST INFRASTRUCTURE (build container only; the fixture is committed).  Weights: synthetic_state_dict(SMALL, 0);
torch.manual_seed / random.seed fix x_T, the per-step noise and the window shifts.

    python oracle/make_golden_loops.py [--reference /root/reference]
"""
from __future__ import annotations

import argparse
import os
import random
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle.make_golden import SMALL, ShiftLog, import_reference, reference_flags  # noqa: E402
from oracle.mmdiff_oracle import synthetic_state_dict  # noqa: E402

CASES = {   # name -> (timestep_respacing, loop, kwargs)
    "ancestral_5": ("5", "p_sample_loop", dict(clip_denoised=True)),
    "ancestral_3_noclip": ("3", "p_sample_loop", dict(clip_denoised=False)),
    "ddim_6_eta0": ("ddim6", "ddim_sample_loop", dict(clip_denoised=True, eta=0.0)),
    "ddim_4_eta05": ("ddim4", "ddim_sample_loop", dict(clip_denoised=True, eta=0.5)),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "loops_small.pt"))
    args = ap.parse_args()
    su = import_reference(args.reference)
    torch.set_num_threads(os.cpu_count() or 1)
    B = 2
    shape = {"video": (B, *SMALL.video_size), "audio": (B, *SMALL.audio_size)}
    fixture = {"config": SMALL.__dict__, "weight_seed": 0, "batch": B, "cases": {}}
    sd = synthetic_state_dict(SMALL, seed=0)
    for name, (respacing, loop, kw) in CASES.items():
        flags = reference_flags(su, SMALL)
        flags["timestep_respacing"] = respacing
        model, diffusion = su.create_model_and_diffusion(**flags)
        model.load_state_dict(sd, strict=True)
        model.eval()
        torch.manual_seed(11)
        random.seed(17)
        with torch.no_grad(), ShiftLog() as log:
            out = getattr(diffusion, loop)(model, shape=shape, progress=False, **kw)
        fixture["cases"][name] = {"respacing": respacing, "loop": loop, "kwargs": kw, "torch_seed": 11, "shift_seed": 17,
                                  "n_draws": len(log.draws), "timestep_map": list(diffusion.timestep_map),
                                  "video": out["video"].clone(), "audio": out["audio"].clone()}
        print(f"{name:20s} steps={diffusion.num_timesteps} |v|={out['video'].norm().item():.4f} |a|={out['audio'].norm().item():.4f}")
    torch.save(fixture, args.out)
    print(f"wrote {args.out} ({os.path.getsize(args.out) / 1e3:.1f} kB)")


if __name__ == "__main__":
    main()
