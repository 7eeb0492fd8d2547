"""Generate tests/golden/*.pt by running the UNMODIFIED reference (imported from /root/reference).

TEST INFRASTRUCTURE.  Runs only in the build container (the reference cannot travel to the GPU
box); the resulting small fixtures are committed.  Weights are not stored: they come from
oracle.mmdiff_oracle.synthetic_state_dict(cfg, seed) and are loaded into the reference model with
load_state_dict, so any checkout can regenerate them bit-exactly.

    python oracle/make_golden.py [--reference /root/reference] [--skip-production]
"""
from __future__ import annotations

import argparse
import os
import random
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle.mmdiff_oracle import UNetConfig, synthetic_state_dict  # noqa: E402

SMALL = UNetConfig(video_size=(8, 3, 16, 16), audio_size=(1, 2048), model_channels=64, num_res_blocks=1,
                   channel_mult=(1, 1, 2), num_heads=1, num_head_channels=64,
                   cross_attention_resolutions=(1, 2, 4), cross_attention_windows=(1, 4, 8),
                   cross_attention_shift=True, video_attention_resolutions=(2, 4),
                   audio_attention_resolutions=(-1,))
PRODUCTION = UNetConfig()


def import_reference(path):
    m = types.ModuleType("mpi4py")
    m.MPI = types.SimpleNamespace(COMM_WORLD=None)
    sys.modules["mpi4py"] = m
    sys.modules["blobfile"] = types.ModuleType("blobfile")
    sys.path.insert(0, path)
    from mm_diffusion import multimodal_script_util as su  # noqa
    return su


def reference_flags(su, cfg: UNetConfig):
    d = su.model_and_diffusion_defaults()
    join = lambda xs: ",".join(str(int(x)) for x in xs)
    d.update(video_size=list(cfg.video_size), audio_size=list(cfg.audio_size), num_channels=cfg.model_channels,
             num_res_blocks=cfg.num_res_blocks, channel_mult=join(cfg.channel_mult), num_heads=cfg.num_heads,
             num_head_channels=cfg.num_head_channels,
             cross_attention_resolutions=join(cfg.cross_attention_resolutions),
             cross_attention_windows=join(cfg.cross_attention_windows),
             cross_attention_shift=cfg.cross_attention_shift,
             video_attention_resolutions=join(cfg.video_attention_resolutions),
             audio_attention_resolutions=join(cfg.audio_attention_resolutions),
             resblock_updown=True, use_scale_shift_norm=True, learn_sigma=False, use_fp16=False, dropout=0.0)
    return d


class ShiftLog:
    """Records the (lo, hi, value) of every random.randint draw the reference makes."""

    def __init__(self):
        self.draws = []
        self._orig = random.randint

    def __enter__(self):
        def rec(lo, hi):
            v = self._orig(lo, hi)
            self.draws.append((lo, hi, v))
            return v
        random.randint = rec
        return self

    def __exit__(self, *a):
        random.randint = self._orig


def make_inputs(cfg, batch, seed):
    g = torch.Generator().manual_seed(seed)
    v = torch.randn(batch, *cfg.video_size, generator=g)
    a = torch.randn(batch, *cfg.audio_size, generator=g)
    return v, a


def golden_for(su, cfg: UNetConfig, name: str, batch: int, weight_seed: int, full: bool, out_dir: str):
    model, diffusion = su.create_model_and_diffusion(**reference_flags(su, cfg))
    sd = synthetic_state_dict(cfg, seed=weight_seed)
    ref_keys = [(k, tuple(v.shape)) for k, v in model.state_dict().items()]
    model.load_state_dict(sd, strict=True)
    model.eval()
    fixture = {"config": cfg.__dict__, "weight_seed": weight_seed, "batch": batch,
               "state_dict_keys": ref_keys,
               "num_params": sum(p.numel() for p in model.parameters())}
    v, a = make_inputs(cfg, batch, seed=1234)
    t = torch.tensor([500, 17][:batch] if batch <= 2 else [500] * batch)
    fixture["t"] = t
    fixture["input_seed"] = 1234
    with torch.no_grad():
        random.seed(7)
        with ShiftLog() as log:
            ev, ea = model(v, a, t)
        fixture["forward_shift_draws"] = log.draws
        if full:
            fixture["forward_video"] = ev.clone()
            fixture["forward_audio"] = ea.clone()
        else:  # production: strided subsample + norms keep the fixture small
            fixture["forward_video_sub"] = ev.flatten()[::8].clone()
            fixture["forward_audio_sub"] = ea.flatten()[::8].clone()
        fixture["forward_video_norm"] = ev.norm().item()
        fixture["forward_audio_norm"] = ea.norm().item()

    if full:
        # --- one p_sample step with injected noise (reference draws th.randn_like twice: video, audio)
        g = torch.Generator().manual_seed(99)
        zv = torch.randn(v.shape, generator=g)
        za = torch.randn(a.shape, generator=g)
        queue = [zv, za]
        orig = torch.randn_like
        torch.randn_like = lambda x, *args, **kw: queue.pop(0)
        try:
            for tag, tt in (("mid", torch.tensor([500, 17][:batch])), ("zero", torch.tensor([0, 999][:batch]))):
                queue[:] = [zv, za]
                random.seed(11)
                with torch.no_grad(), ShiftLog() as log:
                    out = diffusion.p_sample(model, {"video": v, "audio": a}, tt, clip_denoised=True)
                fixture[f"p_sample_{tag}"] = {
                    "t": tt, "shift_draws": log.draws,
                    "sample_video": out["sample"]["video"].clone(), "sample_audio": out["sample"]["audio"].clone(),
                    "pred_start_video": out["pred_start"]["video"].clone(),
                    "pred_start_audio": out["pred_start"]["audio"].clone()}
        finally:
            torch.randn_like = orig
        fixture["noise_seed"] = 99
        # --- training loss (forward only is enough to pin q_sample + loss arithmetic)
        x0v, x0a = v.clamp(-1, 1), a.clamp(-1, 1)
        tt = torch.tensor([321, 900][:batch])
        random.seed(13)
        with torch.no_grad(), ShiftLog() as log:
            losses = diffusion.multimodal_training_losses(model, {"video": x0v, "audio": x0a}, tt,
                                                         noise={"video": zv, "audio": za})
        fixture["training"] = {"t": tt, "shift_draws": log.draws,
                               **{k: val.clone() for k, val in losses.items() if torch.is_tensor(val)}}
        # --- schedule tables
        fixture["alphas_cumprod"] = torch.from_numpy(diffusion.alphas_cumprod.copy())
        fixture["posterior_mean_coef1"] = torch.from_numpy(diffusion.posterior_mean_coef1.copy())
    path = os.path.join(out_dir, f"{name}.pt")
    torch.save(fixture, path)
    print(f"wrote {path} ({os.path.getsize(path) / 1e3:.1f} kB)  |ev|={fixture['forward_video_norm']:.4f} "
          f"|ea|={fixture['forward_audio_norm']:.4f} shifts={[d[2] for d in fixture['forward_shift_draws']]}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    ap.add_argument("--skip-production", action="store_true")
    args = ap.parse_args()
    su = import_reference(args.reference)
    torch.set_num_threads(os.cpu_count() or 1)
    os.makedirs(args.out, exist_ok=True)
    golden_for(su, SMALL, "small", batch=2, weight_seed=0, full=True, out_dir=args.out)
    if not args.skip_production:
        golden_for(su, PRODUCTION, "production", batch=1, weight_seed=0, full=False, out_dir=args.out)


if __name__ == "__main__":
    main()
