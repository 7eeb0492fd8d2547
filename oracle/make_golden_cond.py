"""Generate tests/golden/cond_small.pt by running the UNMODIFIED reference zero-shot conditional sampling loops
(/root/reference/mm_diffusion/multimodal_gaussian_diffusion.py:584-819: replacement method class_scale = 0 and
gradient guidance class_scale > 0, the latter differentiating through the reference model) on the SMALL reference
model, CPU fp32, with 4 respaced steps.

TEST INFRASTRUCTURE (build container only; the fixture is committed).  Weights: synthetic_state_dict(SMALL, 0);
torch.manual_seed / random.seed fix x_T, the per-step noise and the window shifts; the condition is 0.1 * randn.

    python oracle/make_golden_cond.py [--reference /root/reference]
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

CASES = {   # name -> (conditioned modality, class_scale, use_fp16 flag of the loop (only scales the guidance loss))
    "a2v_replace": ("audio", 0.0, False),
    "v2a_replace": ("video", 0.0, False),
    "a2v_grad": ("audio", 3.0, False),
    "v2a_grad": ("video", 1.5, False),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "cond_small.pt"))
    args = ap.parse_args()
    su = import_reference(args.reference)
    torch.set_num_threads(os.cpu_count() or 1)
    flags = reference_flags(su, SMALL)
    flags["timestep_respacing"] = "4"
    model, diffusion = su.create_model_and_diffusion(**flags)
    model.load_state_dict(synthetic_state_dict(SMALL, seed=0), strict=True)
    model.eval()
    B = 2
    shape = {"video": (B, *SMALL.video_size), "audio": (B, *SMALL.audio_size)}
    g = torch.Generator().manual_seed(77)
    cond = {"video": 0.1 * torch.randn(B, *SMALL.video_size, generator=g), "audio": 0.1 * torch.randn(B, *SMALL.audio_size, generator=g)}
    fixture = {"config": SMALL.__dict__, "weight_seed": 0, "batch": B, "cond_seed": 77, "respacing": "4", "cases": {}}
    for name, (modality, scale, fp16) in CASES.items():
        torch.manual_seed(5)
        random.seed(13)
        # Gradient cases: the reference re-runs every (always checkpointed) CrossAttentionBlock in backward and DRAWS A NEW
        # window shift there (multimodal_unet.py:619-622 under nn.py:233-279), i.e. it differentiates a different function
        # than it evaluated.  The fixture pins the draws to a constant so forward and recomputation agree and the
        # gradient is the gradient of the evaluated function (what the sm_100a backward computes).
        orig_randint = random.randint
        if scale != 0:
            random.randint = lambda lo, hi: min(hi, 3)
        try:
            with ShiftLog() as log:
                out = diffusion.conditional_p_sample_loop(model, shape=shape, use_fp16=fp16, clip_denoised=True,
                                                          model_kwargs={modality: cond[modality].clone()}, progress=False,
                                                          class_scale=scale)
        finally:
            random.randint = orig_randint
        fixture["cases"][name] = {"condition": modality, "class_scale": scale, "use_fp16": fp16, "torch_seed": 5, "shift_seed": 13,
                                  "shift_draws": list(log.draws), "const_shift": 3 if scale != 0 else None, "video": out["video"].detach().clone(),
                                  "audio": out["audio"].detach().clone()}
        print(f"{name:12s} |v|={out['video'].norm().item():.4f} |a|={out['audio'].norm().item():.4f} draws={len(log.draws)}")
    torch.save(fixture, args.out)
    print(f"wrote {args.out} ({os.path.getsize(args.out) / 1e3:.1f} kB)")


if __name__ == "__main__":
    main()
