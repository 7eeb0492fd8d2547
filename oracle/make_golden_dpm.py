"""Generate tests/golden/dpm_small.pt by running the UNMODIFIED reference DPM_Solver
(/root/reference/mm_diffusion/multimodal_dpm_solver_plus.py) on the SMALL reference model, CPU fp32.

TEST INFRASTRUCTURE (build container only; the fixture is committed).  Weights come from
oracle.mmdiff_oracle.synthetic_state_dict(SMALL, 0); x_T from torch.Generator(1234); the cross-attention shifts the
reference draws from the global `random` are seeded per case and recorded, as are the integer model times of every
evaluation.

    python oracle/make_golden_dpm.py [--reference /root/reference]
"""
from __future__ import annotations

import argparse
import contextlib
import io
import os
import random
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle.make_golden import SMALL, ShiftLog, import_reference, make_inputs, reference_flags  # noqa: E402
from oracle.mmdiff_oracle import synthetic_state_dict  # noqa: E402

# name -> (DPM_Solver kwargs, sample kwargs).  Shortened versions of the reference's call sites:
#   multimodal_sample_sr.py:125-131 (steps=20, order=3, logSNR, singlestep), :139-145 (dpm_solver++: predict_x0 +
#   thresholding, order=2, logSNR, adaptive), BASELINE configs[2] (multistep order 2, time_uniform, predict_x0 + thresholding).
CASES = {
    "ms2_x0_thr_tu": (dict(predict_x0=True, thresholding=True), dict(steps=8, order=2, skip_type="time_uniform", method="multistep")),
    "ss3_eps_logsnr": (dict(), dict(steps=7, order=3, skip_type="logSNR", method="singlestep")),
    "ss3_x0_logsnr": (dict(predict_x0=True), dict(steps=6, order=3, skip_type="logSNR", method="singlestep")),
    "ss2_x0_tq": (dict(predict_x0=True), dict(steps=5, order=2, skip_type="time_quadratic", method="singlestep")),
    "ssfixed2_eps_tu": (dict(), dict(steps=6, order=2, skip_type="time_uniform", method="singlestep_fixed")),
    "ms1_eps_tu": (dict(), dict(steps=4, order=1, skip_type="time_uniform", method="multistep")),
    "ms2_eps_logsnr": (dict(), dict(steps=6, order=2, skip_type="logSNR", method="multistep")),
    "ms2_x0_taylor_denoise": (dict(predict_x0=True), dict(steps=5, order=2, skip_type="time_uniform", method="multistep",
                                                         solver_type="taylor", denoise=True)),
    "ss2_eps_taylor": (dict(), dict(steps=4, order=2, skip_type="time_uniform", method="singlestep", solver_type="taylor")),
    "adaptive2_x0_thr": (dict(predict_x0=True, thresholding=True), dict(order=2, skip_type="logSNR", method="adaptive")),
    "adaptive3_eps": (dict(), dict(order=3, method="adaptive", atol=0.05, rtol=0.2)),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "dpm_small.pt"))
    args = ap.parse_args()
    su = import_reference(args.reference)
    from mm_diffusion.multimodal_dpm_solver_plus import DPM_Solver, NoiseScheduleVP
    torch.set_num_threads(os.cpu_count() or 1)
    model, diffusion = su.create_model_and_diffusion(**reference_flags(su, SMALL))
    model.load_state_dict(synthetic_state_dict(SMALL, seed=0), strict=True)
    model.eval()
    acp = torch.tensor(diffusion.alphas_cumprod, dtype=torch.float32)

    times = []

    class Recorder(torch.nn.Module):
        """Passes through to the reference model, recording the integer times it is evaluated at."""
        def __init__(self, m):
            super().__init__()
            self.m = m
            self.video_out_channels = m.video_out_channels
            self.audio_out_channels = m.audio_out_channels

        def forward(self, video, audio, t, **kw):
            times.append(t.clone())
            return self.m(video, audio, t, **kw)

    rec = Recorder(model)
    batch = 2
    xv, xa = make_inputs(SMALL, batch, seed=1234)
    fixture = {"config": SMALL.__dict__, "weight_seed": 0, "batch": batch, "input_seed": 1234, "cases": {}}

    # schedule pins: lambda, its inverse, sigma on a grid of times (incl. grid points and out-of-range values)
    ns = NoiseScheduleVP("discrete", alphas_cumprod=acp)
    tq = torch.cat([torch.linspace(1e-3, 1.0, 37), torch.tensor([0.0005, 0.00137, 0.5, 0.9995, 1.0])])
    lam = ns.marginal_lambda(tq)
    fixture["schedule"] = {"t": tq, "log_alpha": ns.marginal_log_mean_coeff(tq), "sigma": ns.marginal_std(tq), "lambda": lam,
                           "inverse_lambda": ns.inverse_lambda(lam)}

    for name, (ckw, skw) in CASES.items():
        times.clear()
        solver = DPM_Solver(model=rec, alphas_cumprod=acp, **ckw)
        random.seed(21)
        sink = io.StringIO()
        with torch.no_grad(), ShiftLog() as log, contextlib.redirect_stdout(sink):
            out = solver.sample({"video": xv.clone(), "audio": xa.clone()}, **skw)
        fixture["cases"][name] = {
            "solver_kwargs": ckw, "sample_kwargs": skw, "shift_seed": 21, "shift_draws": list(log.draws),
            "model_times": torch.stack([t.to(torch.int64) for t in times]),   # [NFE, B]
            "video": out["video"].clone(), "audio": out["audio"].clone()}
        print(f"{name:24s} NFE={len(times):3d} |v|={out['video'].norm().item():.4f} |a|={out['audio'].norm().item():.4f} "
              f"t[:6]={[int(t[0]) for t in times[:6]]}")
    torch.save(fixture, args.out)
    print(f"wrote {args.out} ({os.path.getsize(args.out) / 1e3:.1f} kB)")


if __name__ == "__main__":
    main()
