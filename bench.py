#!/usr/bin/env python
"""Benchmark of the MM-Diffusion denoising hot path on B200 (contract: see task prompt / DESIGN.md §measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A "step" = one p_sample step (MultimodalUNet.forward + sampler tail) over one batch of synthetic input.
Workload at N=1 = BASELINE.json configs[1]: the per-step unit of the 1000-step DDPM p_sample_loop at batch 4,
16x3x64x64 video + 1x25600 audio, random-init production U-Net (133.7 M params).  N>1: one process per GPU
(torchrun), batch 4 per rank (weak scaling), no collective inside a step; the finished samples are gathered
once with NCCL at the end of the timed region.  Metric: denoising steps/sec = samples x steps / time.

Other workloads (BASELINE.json configs[2..4]) ride along as `extra` sub-records of the default line and can be run
alone with --workload {dpm,cond,train}:
  dpm   DPM-Solver++ (predict_x0, thresholding) 50 NFE, order 2, time_uniform, multistep, batch 16 per GPU   (configs[2])
  train multimodal_training_losses forward + backward, batch 8 per GPU, one flat gradient all-reduce          (configs[3])
  cond  audio -> video replacement-conditioned ancestral step (class_scale 0), batch 4 per GPU                (configs[4])

--impl reference times the reference's own CPU implementation of the path on the host cores for the same metric and
config: the UNMODIFIED reference from baseline/_ref (tools/install_reference.py; kind "reference"), or the oracle port
(oracle/mmdiff_oracle.py; kind "port") when that tree is absent.  The default line also carries `gpu_eager_baseline`:
the unmodified reference run through PyTorch eager on the same GPU (fp16 production flags and fp32) — none of this
repository's modules are on that path.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "denoising steps/sec (16fx64x64 video + 25600 audio)"
UNIT = "sample-steps/s"
VIDEO_SIZE = [16, 3, 64, 64]
AUDIO_SIZE = [1, 25600]


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p.get("hbm_gbs", 6650.0), "bf16_tflops": p.get("bf16_tflops", 1590.0),
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", 1400.0), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []
        self.begin = 0

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self):
        """Start of the timed region: nvidia-smi is started earlier (its first sample takes a few hundred ms, longer than a
        short timed region), only the samples from here on count."""
        self.begin = len(self.lines)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        if len(self.lines) <= self.begin:   # nothing inside the region yet: take the next sample (flagged below)
            t_end = time.time() + 1.0
            while not self.lines and time.time() < t_end:
                time.sleep(0.02)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        window = self.lines[self.begin:]
        outside = False
        if not window and self.lines:   # region shorter than one sampling period: the sample taken right before it
            window, outside = self.lines[-1:], True
        for ln in window:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); smax.append(float(parts[2])); power.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "power_w_max": max(power),
               "samples": len(sm), "reasons": sorted(reasons)}
        if outside:
            out["note"] = "no nvidia-smi sample fell inside the timed region (100 ms period): nearest sample outside it"
        return out


def production_flags():
    from mm_diffusion_b200.script_util import model_and_diffusion_defaults
    d = model_and_diffusion_defaults()
    d.update(video_size=VIDEO_SIZE, audio_size=AUDIO_SIZE, num_channels=128, num_res_blocks=2, num_head_channels=64,
             cross_attention_resolutions="2,4,8", cross_attention_windows="1,4,8", cross_attention_shift=True,
             video_attention_resolutions="2,4,8", audio_attention_resolutions="-1", resblock_updown=True,
             use_scale_shift_norm=True, learn_sigma=False, use_fp16=True, diffusion_steps=1000, noise_schedule="linear")
    return d


def build_b200(device, seed=0):
    """Random-init production model; the reference's zero-initialised tensors are re-drawn N(0, 0.02^2) so no branch
    is dead (timing is weight independent)."""
    import torch
    from mm_diffusion_b200.script_util import create_model_and_diffusion
    from mm_diffusion_b200.unet import _ZERO_INIT_MARKERS
    torch.manual_seed(seed)
    model, diffusion = create_model_and_diffusion(**production_flags())
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if any(mk in name for mk in _ZERO_INIT_MARKERS):
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
    model.to(device).eval()
    model.convert_to_fp16()
    return model, diffusion


# ----------------------------------------------------------------------------- the unmodified reference (baseline/_ref)
REF_ROOT = os.path.join(ROOT, "baseline", "_ref")
_REF_ZERO_MARKERS = (".video_out_layers.3.", ".audio_out_layers.3.", ".proj_out.", ".video_proj_out.", ".audio_proj_out.",
                     "video_out.2.", "audio_out.2.")


def import_reference():
    """mm_diffusion.multimodal_script_util of the UNMODIFIED reference (None when baseline/_ref is absent).  Only the
    import-time stand-ins for mpi4py / blobfile (tools/ref_shims) are added; no module of this repository is involved."""
    if not os.path.isdir(os.path.join(REF_ROOT, "mm_diffusion")):
        return None
    for p in (REF_ROOT, os.path.join(ROOT, "tools", "ref_shims")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from mm_diffusion import multimodal_script_util as su
    return su


def build_reference(su, device, use_fp16, seed=0):
    """Reference production model through its own factory (multimodal_script_util.py:62-128), same synthetic weights
    policy as build_b200 (zero-initialised tensors re-drawn so no branch is dead)."""
    import torch
    d = su.model_and_diffusion_defaults()
    d.update(video_size=VIDEO_SIZE, audio_size=AUDIO_SIZE, num_channels=128, num_res_blocks=2, num_head_channels=64,
             cross_attention_resolutions="2,4,8", cross_attention_windows="1,4,8", cross_attention_shift=True,
             video_attention_resolutions="2,4,8", audio_attention_resolutions="-1", resblock_updown=True,
             use_scale_shift_norm=True, learn_sigma=False, use_fp16=use_fp16, diffusion_steps=1000, noise_schedule="linear")
    torch.manual_seed(seed)
    model, diffusion = su.create_model_and_diffusion(**d)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if any(mk in name for mk in _REF_ZERO_MARKERS):
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
    model.to(device).eval()
    if use_fp16:
        model.convert_to_fp16()
    return model, diffusion


def gpu_eager_baseline(device, batch, warmup=2, steps=5):
    """The unmodified reference through PyTorch eager on this GPU: create_model_and_diffusion + p_sample
    at the bench batch, fp16 (the production flags) and fp32, CUDA events around `steps` p_sample calls."""
    import torch
    su = import_reference()
    if su is None:
        return {"unavailable": "baseline/_ref absent (python tools/install_reference.py)"}
    out = {"how": f"unmodified reference, PyTorch {torch.__version__} eager, batch {batch}, {warmup} warm-up + {steps} timed "
                  "p_sample steps, CUDA events; TF32 settings left at the PyTorch defaults"}
    for tag, fp16 in (("fp16", True), ("fp32", False)):
        try:
            model, diffusion = build_reference(su, device, fp16)
            g = torch.Generator().manual_seed(1234)
            x = {"video": torch.randn(batch, *VIDEO_SIZE, generator=g).to(device),
                 "audio": torch.randn(batch, *AUDIO_SIZE, generator=g).to(device)}
            t = torch.full((batch,), 500, device=device, dtype=torch.long)
            with torch.no_grad():
                for _ in range(warmup):
                    x = diffusion.p_sample(model, x, t)["sample"]
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(steps):
                    x = diffusion.p_sample(model, x, t)["sample"]
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[tag] = {"value": round(batch / (ms * 1e-3), 3), "unit": UNIT, "ms_per_step": round(ms, 3), "steps": steps,
                        "finite": bool(torch.isfinite(x["video"].float()).all().item())}
            del model, diffusion, x
            torch.cuda.empty_cache()
        except Exception as e:  # noqa: BLE001  (a baseline that cannot run is reported, never fatal for the bench line)
            out[tag] = {"error": f"{type(e).__name__}: {str(e)[:300]}"}
    return out


# step family (model.cu StepInfo.kind) -> kernel function that executes it
KERNEL_OF = {"conv3x3_spatial": "conv_gemm_kernel", "conv1x1_qkv": "conv_gemm_kernel", "conv1x1_out": "conv_gemm_kernel",
             "conv1x1_proj": "conv_gemm_kernel", "conv_temporal": "conv_gemm_kernel", "conv_audio_k3": "conv_gemm_kernel",
             "conv_head": "conv_gemm_kernel", "conv_stem": "conv_gemm_kernel",
             "cross_attention": "attention64_kernel", "self_attention": "attention64_kernel",
             "group_norm": "gn_apply_kernel+gn_stats_kernel", "temporal_attention": "temporal_attn_kernel",
             "resample": "resample_kernel", "im2col": "im2col_kernel", "time_embed": "time_embed_kernel"}


def ncu_traffic(kernel, batch):
    """Measured DRAM bytes per launch of `kernel` (ncu dram__bytes_read.sum + dram__bytes_write.sum averaged over the
    launches of one forward at this batch), from the committed summary of the capture; None if not captured."""
    import glob
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_dram_traffic.json")), reverse=True):   # latest round first
        try:
            with open(path) as f:
                t = json.load(f)
            if t.get("batch") != batch:
                continue
            ks = t["kernels"]
            hit = [v for k, v in ks.items() if kernel in k]
            if hit:
                n = sum(v.get("launches", 1) for v in hit)
                return sum(v["dram_bytes_per_launch"] * v.get("launches", 1) for v in hit) / max(n, 1), os.path.basename(path)
        except Exception:
            continue
    return None


def family_summary(steps):
    fam = {}
    for s in steps:
        f = fam.setdefault(s["kind"], {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
        f["ms"] += s["ms"]; f["flops"] += s["flops"]; f["bytes"] += s["bytes"]; f["launches"] += s["kernels"]
    return fam


class Ctx:
    """Process-group context of one bench process (one process per GPU)."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device — the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local_rank)
        self.device = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.device)

    def barrier(self):
        import torch
        import torch.distributed as dist
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, *vals):
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return vals
        t = torch.tensor(list(vals), device=self.device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return tuple(t.tolist())

    def close(self):
        import torch.distributed as dist
        if self.world > 1:
            dist.barrier()
            dist.destroy_process_group()


def timed(ctx, step, warm, k, clock_sampler=None):
    """`warm` untimed + EXACTLY `k` timed calls of step(i), bracketed by barrier + synchronize, CUDA events on the
    current stream; returns the device milliseconds of the timed region (this rank)."""
    import torch
    if clock_sampler is not None:
        clock_sampler.start()
    for i in range(warm):
        step(i)
    ctx.barrier()
    if clock_sampler is not None:
        clock_sampler.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(k):
        step(warm + i)
    e1.record()
    ctx.barrier()
    return e0.elapsed_time(e1)


def run_b200(ctx, args, model, diffusion):
    import torch
    import torch.distributed as dist
    world, rank, local_rank, device = ctx.world, ctx.rank, ctx.local_rank, ctx.device
    B, K, W = args.batch, args.steps, max(args.warmup, 3)
    import random
    random.seed(4321 + rank)
    gen = torch.Generator().manual_seed(1234 + rank)
    xv_h = torch.randn(B, *VIDEO_SIZE, generator=gen).pin_memory()
    xa_h = torch.randn(B, *AUDIO_SIZE, generator=gen).pin_memory()
    torch.manual_seed(99 + rank)
    T = diffusion.num_timesteps
    ts = [torch.full((B,), (T - 1 - i) % T, device=device, dtype=torch.long) for i in range(W + K)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def sample_epilogue(x):
        """End-of-loop work of the sample scripts: uint8 video, and (N>1) the one NCCL gather of finished samples."""
        from mm_diffusion_b200.parallel import sample_epilogue as fused_epilogue
        v8 = fused_epilogue(x["video"])   # uint8 + channels-last permute of the script (:159-163) in one kernel
        a = x["audio"]
        if world > 1:
            gv = torch.empty((world,) + tuple(v8.shape), dtype=torch.uint8, device=device)
            ga = torch.empty((world,) + tuple(a.shape), dtype=a.dtype, device=device)
            dist.all_gather_into_tensor(gv, v8.contiguous())
            dist.all_gather_into_tensor(ga, a.contiguous())
            return gv, ga
        return v8, a

    # ---------------- device-resident loop (the headline `value`)
    x = {"video": xv_h.to(device), "audio": xa_h.to(device)}
    with torch.no_grad():
        clocks = ClockSampler(local_rank)
        if rank == 0:
            clocks.start()   # nvidia-smi needs a few hundred ms to its first sample: started before the warm-up
        for i in range(W):
            x = diffusion.p_sample(model, x, ts[i])["sample"]
        sample_epilogue(x)   # warm the end-of-loop path too (first NCCL call builds the communicator)
        barrier()
        clocks.mark()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K):
            x = diffusion.p_sample(model, x, ts[W + i])["sample"]
        out = sample_epilogue(x)
        e1.record()
        barrier()
        clk = clocks.stop() if rank == 0 else None
    ms_total = e0.elapsed_time(e1)
    finite = bool(torch.isfinite(x["video"]).all().item() and torch.isfinite(x["audio"]).all().item())

    # ---------------- end-to-end through the public API with host buffers
    hv_out = torch.empty(B, *VIDEO_SIZE).pin_memory()
    ha_out = torch.empty(B, *AUDIO_SIZE).pin_memory()
    with torch.no_grad():
        def e2e_step(i):
            xd = {"video": xv_h.to(device, non_blocking=True), "audio": xa_h.to(device, non_blocking=True)}
            s = diffusion.p_sample(model, xd, ts[i])["sample"]
            hv_out.copy_(s["video"], non_blocking=True)
            ha_out.copy_(s["audio"], non_blocking=True)
        for i in range(W):
            e2e_step(i)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for i in range(K):
            e2e_step(W + i)
        f1.record()
        barrier()
    ms_e2e = f0.elapsed_time(f1)

    if world > 1:
        tmax = torch.tensor([ms_total, ms_e2e], device=device)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms_total, ms_e2e = tmax[0].item(), tmax[1].item()

    launches_fwd = model.num_launches(B)
    result = None
    if rank == 0:
        peaks = load_peaks()
        # per-launch device times of one forward (un-graphed, CUDA events on the launch stream)
        steps = model.profile(B, reps=args.profile_reps)
        fam = family_summary(steps)
        fwd_ms = sum(s["ms"] for s in steps)
        # group the step families by the kernel function that runs them; the roofline is the dominant kernel's
        kern = {}
        for k, v in fam.items():
            kk = kern.setdefault(KERNEL_OF.get(k, k), {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
            for f in ("ms", "flops", "bytes", "launches"):
                kk[f] += v[f]
        dname, d = max(((k, v) for k, v in kern.items() if v["launches"] > 0), key=lambda kv: kv[1]["ms"])
        tflops = d["flops"] / (d["ms"] * 1e-3) / 1e12 if d["ms"] > 0 else 0.0
        gbs = d["bytes"] / (d["ms"] * 1e-3) / 1e9 if d["ms"] > 0 else 0.0
        ridge = (peaks["bf16_tflops_sustained"] * 1e12) / (peaks["hbm_gbs"] * 1e9)
        tensor_bound = (d["flops"] / max(d["bytes"], 1.0)) > ridge
        if tensor_bound:
            roof = {"bound": "tensor", "achieved": round(tflops, 2), "peak": peaks["bf16_tflops_sustained"],
                    "unit": "TFLOP/s", "frac": round(tflops / peaks["bf16_tflops_sustained"], 4)}
        else:
            roof = {"bound": "hbm", "achieved": round(gbs, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": round(gbs / peaks["hbm_gbs"], 4)}
        traffic_src = ncu_traffic(dname, B)
        traffic, traffic_file = traffic_src if traffic_src else (None, None)
        roof.update({"kernel": dname, "launches_per_step": d["launches"], "share_of_step": round(d["ms"] / fwd_ms, 4),
                     "algorithmic_gflop_per_launch": round(d["flops"] / d["launches"] / 1e9, 3),
                     "algorithmic_mb_per_launch": round(d["bytes"] / d["launches"] / 1e6, 3),
                     "us_per_launch": round(d["ms"] * 1e3 / d["launches"], 2),
                     "intensity_flop_per_byte": round(d["flops"] / max(d["bytes"], 1.0), 1), "ridge_flop_per_byte": round(ridge, 1),
                     "hbm_gbs": round(gbs, 1), "hbm_frac": round(gbs / peaks["hbm_gbs"], 4),
                     "peak_source": peaks["source"] + (" (sustained cuBLAS bf16 GEMM; fp16 runs at the same rate)"
                                                       if tensor_bound else " (copy)"),
                     "traffic": traffic,
                     "traffic_note": "mean dram__bytes_read+write per launch of this kernel over one forward, ncu capture "
                                     f"summarised in profiles/{traffic_file}" if traffic is not None else None,
                     "how": "sum of the algorithmic FLOPs/bytes of this kernel's launches in one forward / sum of their "
                            f"CUDA-event durations on the launch stream, mean of {args.profile_reps} un-graphed passes "
                            "after the timed region"})
        kernels = {k: {"ms": round(v["ms"], 4), "launches": v["launches"],
                       "tflops": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 2) if v["ms"] > 0 else 0.0,
                       "gbs": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1) if v["ms"] > 0 else 0.0}
                   for k, v in sorted(kern.items(), key=lambda kv: -kv[1]["ms"]) if v["launches"] > 0}
        families = {k: {"ms": round(v["ms"], 4), "launches": v["launches"],
                        "tflops": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 2) if v["ms"] > 0 else 0.0,
                        "gbs": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1) if v["ms"] > 0 else 0.0}
                    for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])}
        step_flops = sum(s["flops"] for s in steps)
        value = world * B * K / (ms_total * 1e-3)
        e2e_value = world * B * K / (ms_e2e * 1e-3)
        bytes_in = (xv_h.numel() + xa_h.numel()) * 4
        result = {
            "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": round(ms_total / K, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16", "data": "synthetic",
            "config": {"workload": "p_sample step of the 1000-step DDPM loop (BASELINE.json configs[1])",
                       "batch_per_gpu": B, "global_batch": B * world, "video": VIDEO_SIZE, "audio": AUDIO_SIZE,
                       "params_m": round(sum(p.numel() for p in model.parameters()) / 1e6, 2),
                       "parallelism": f"batch-shard x{world} (replicated weights, no in-step collective)",
                       "l2_note": "per-step working set (activations+weights ~2.9 GB at B=4) exceeds the 126 MB L2",
                       "cuda_graph": True},
            "e2e": {"value": round(e2e_value, 3), "unit": UNIT, "ms_per_step": round(ms_e2e / K, 4),
                    "h2d_bytes_per_step": bytes_in, "d2h_bytes_per_step": bytes_in},
            "gpu_launches": K * (launches_fwd + 3),
            "launches_per_step": launches_fwd + 3,
            "model_tflops": round(step_flops * K / (ms_total * 1e-3) / 1e12, 2),
            "forward_ms_ungraphed": round(fwd_ms, 3),
            "roofline": roof, "kernels": kernels, "families": families, "clocks": clk, "finite": finite,
        }
        result["warmup_requested"] = args.warmup   # at least 3 warm-up steps are always run (timing rules)
        if world == 1 and not args.no_cpu_baseline:
            result["cpu_baseline"] = cpu_baseline(B, max_seconds=45.0)
        if world == 1 and not args.no_gpu_baseline:
            result["gpu_eager_baseline"] = gpu_eager_baseline(device, B)
    return result


def _common_line(ctx, metric, value, K, W, ms_total, workload, B, extra_cfg=None):
    cfg = {"workload": workload, "batch_per_gpu": B, "global_batch": B * ctx.world, "video": VIDEO_SIZE, "audio": AUDIO_SIZE,
           "cuda_graph": True}
    cfg.update(extra_cfg or {})
    return {"metric": metric, "value": round(value, 3), "unit": UNIT, "n_gpus": ctx.world, "steps": K, "warmup": W,
            "ms_per_step": round(ms_total / K, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16", "data": "synthetic", "config": cfg}


def _tensor_roof(flops_per_step, ms_per_step, what):
    peaks = load_peaks()
    tf = flops_per_step / (ms_per_step * 1e-3) / 1e12
    return {"bound": "tensor", "achieved": round(tf, 2), "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
            "frac": round(tf / peaks["bf16_tflops_sustained"], 4), "kernel": what, "traffic": None,
            "peak_source": peaks["source"] + " (sustained cuBLAS bf16 GEMM)"}


FLOPS_PER_EVAL = 1.3288e12   # algorithmic FLOPs per model evaluation per sample


def run_b200_train(ctx, args, model, diffusion):
    """BASELINE.json configs[3]: multimodal_training_losses forward + backward, batch 8 per GPU, synthetic
    Landscape-shape data, batch-sharded over the GPUs of the box; the step's one collective is the all-reduce of the
    flat fp32 gradient buffer (or torch DDP's buckets with --ddp).  A step = one forward + backward over one batch (no
    optimizer: the config names fwd+bwd); value = samples x steps / s."""
    import torch
    import torch.distributed as dist
    world, rank, local_rank, device = ctx.world, ctx.rank, ctx.local_rank, ctx.device
    B, K, W = (args.batch if (args.batch_set and args.workload == "train") else 8), args.train_steps or args.steps, max(args.warmup, 3)
    was_dtype = model.dtype
    model.convert_to_fp32()
    model.train()
    model.dropout = args.dropout
    net = model
    use_ddp = world > 1 and args.ddp
    if use_ddp:
        from torch.nn.parallel import DistributedDataParallel as DDP
        net = DDP(model, device_ids=[local_rank], broadcast_buffers=False, bucket_cap_mb=128)
    model.use_flat_gradients(not use_ddp)
    from mm_diffusion_b200.parallel import allreduce_flat_gradients
    import random
    random.seed(4321 + rank)
    gen = torch.Generator().manual_seed(1234 + rank)
    xv_h = torch.randn(B, *VIDEO_SIZE, generator=gen).clamp(-1, 1).pin_memory()
    xa_h = torch.randn(B, *AUDIO_SIZE, generator=gen).clamp(-1, 1).pin_memory()
    t = torch.randint(0, diffusion.num_timesteps, (B,), generator=torch.Generator().manual_seed(99 + rank)).to(device)
    torch.manual_seed(7 + rank)
    comm_ms = []

    def step(x0, measure_comm=False):
        model.zero_grad(set_to_none=True)
        terms = diffusion.multimodal_training_losses(net, x0, t)
        loss = terms["loss"].mean()
        loss.backward()
        if world > 1 and not use_ddp:
            if measure_comm:
                c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                c0.record()
            allreduce_flat_gradients(model)   # the step's one collective: flat fp32 gradient buffer over NCCL
            if measure_comm:
                c1.record()
                comm_ms.append((c0, c1))
        return loss

    x0 = {"video": xv_h.to(device), "audio": xa_h.to(device)}
    clocks = ClockSampler(local_rank) if rank == 0 else None
    last = {}

    def dev_step(i):
        last["loss"] = step(x0)
    ms_total = timed(ctx, dev_step, W, K, clocks)
    clk = clocks.stop() if rank == 0 else None
    loss = last["loss"]
    finite = bool(torch.isfinite(loss).item() and torch.isfinite(model.flat_grad).all().item())

    def e2e_step(i):   # host batch -> device, step, loss back to the host
        step({"video": xv_h.to(device, non_blocking=True), "audio": xa_h.to(device, non_blocking=True)}).item()
    ms_e2e = timed(ctx, e2e_step, W, K)
    exposed = None
    if world > 1 and not use_ddp:   # exposed time of the collective (it is not overlapped with the backward)
        for _ in range(3):
            step(x0, measure_comm=True)
        torch.cuda.synchronize()
        exposed = statistics.median(a.elapsed_time(b) for a, b in comm_ms)
    ms_total, ms_e2e = ctx.max_over_ranks(ms_total, ms_e2e)
    result = None
    if rank == 0:
        fwd_steps = model.plan_steps(B)   # inference plan may be absent; launch counts come from the training plan below
        bwd = model.profile_backward(B, reps=1) if args.profile_reps > 0 else []
        fam = {}
        for s_ in bwd:
            f = fam.setdefault(s_["kind"], {"ms": 0.0, "steps": 0})
            f["ms"] += s_["ms"]; f["steps"] += 1
        step_flops = 3.0 * FLOPS_PER_EVAL * B  # forward + dgrad + wgrad
        value = world * B * K / (ms_total * 1e-3)
        result = _common_line(ctx, "training_losses forward+backward samples/sec (16fx64x64 video + 25600 audio)", value, K, W,
                              ms_total, "multimodal_training_losses forward+backward (BASELINE.json configs[3])", B,
                              {"params_m": round(sum(p.numel() for p in model.parameters()) / 1e6, 2), "dropout": args.dropout,
                               "parallelism": f"batch-shard x{world}" + ((" (torch DDP buckets over NCCL)" if use_ddp else
                                                                          " (one flat fp32 gradient all-reduce over NCCL per step)") if world > 1 else ""),
                               "l2_note": "kept activations + gradients (~48 GB at B=8) exceed the 126 MB L2",
                               "cuda_graph": os.environ.get("MMD_TRAIN_GRAPH", "1") != "0"})
        result.update({
            "e2e": {"value": round(world * B * K / (ms_e2e * 1e-3), 3), "unit": UNIT, "ms_per_step": round(ms_e2e / K, 3),
                    "h2d_bytes_per_step": (xv_h.numel() + xa_h.numel()) * 4, "d2h_bytes_per_step": 4},
            "gpu_launches": K * (model.num_backward_launches(B) + len(fwd_steps)),
            "model_tflops": round(step_flops * K / (ms_total * 1e-3) / 1e12, 2),
            "roofline": _tensor_roof(step_flops, ms_total / K, "whole step (forward + dgrad + wgrad, 3 x 1.3288 TFLOP per sample)"),
            "backward_families": {k: {"ms": round(v["ms"], 3), "steps": v["steps"]} for k, v in
                                  sorted(fam.items(), key=lambda kv: -kv[1]["ms"])},
            "clocks": clk, "finite": finite})
        if exposed is not None:
            result["collective"] = {"what": "NCCL all_reduce of model.flat_grad (fp32, %.2f GB), after the backward" %
                                            (model.flat_grad.numel() * 4 / 1e9), "exposed_ms_per_step": round(exposed, 3)}
    model.zero_grad(set_to_none=True)
    model.use_flat_gradients(False)
    model.dropout = 0
    model.eval()
    if was_dtype != torch.float32:
        model.convert_to_fp16()
    return result


def run_b200_dpm(ctx, args, model, diffusion):
    """BASELINE.json configs[2]: DPM-Solver++ (predict_x0, dynamic thresholding) 50 NFE, order 2, time_uniform, multistep
    (the arguments of the reference's only 50-step call sites, py_scripts/multimodal_sample_sr.py:209-215), batch 16 per
    GPU.  A step = one NFE (model evaluation + solver update) over the batch; whole sample() runs are timed."""
    import torch
    from mm_diffusion_b200.dpm_solver import DPM_Solver
    world, rank, device = ctx.world, ctx.rank, ctx.device
    B = args.batch if (args.batch_set and args.workload == "dpm") else 16
    NFE = 50
    runs = max(1, (args.steps + NFE - 1) // NFE) if args.workload == "dpm" else 1
    import random
    random.seed(4321 + rank)
    gen = torch.Generator().manual_seed(1234 + rank)
    xv_h = torch.randn(B, *VIDEO_SIZE, generator=gen).pin_memory()
    xa_h = torch.randn(B, *AUDIO_SIZE, generator=gen).pin_memory()
    hv_out, ha_out = torch.empty(B, *VIDEO_SIZE).pin_memory(), torch.empty(B, *AUDIO_SIZE).pin_memory()
    solver = DPM_Solver(model, alphas_cumprod=torch.tensor(diffusion.alphas_cumprod, dtype=torch.float32),
                        predict_x0=True, thresholding=True)
    calls = {"n": 0}
    orig_forward = model.forward

    def sample(x):
        with torch.no_grad():
            return solver.sample(x, steps=NFE, order=2, skip_type="time_uniform", method="multistep")
    xd = {"video": xv_h.to(device), "audio": xa_h.to(device)}
    last = {}

    def dev_run(i):
        last["x"] = sample(xd)

    def e2e_run(i):
        x = sample({"video": xv_h.to(device, non_blocking=True), "audio": xa_h.to(device, non_blocking=True)})
        hv_out.copy_(x["video"], non_blocking=True)
        ha_out.copy_(x["audio"], non_blocking=True)
    ms_total = timed(ctx, dev_run, 1, runs)     # warm-up = one whole 50-NFE run
    ms_e2e = timed(ctx, e2e_run, 1, runs)
    finite = bool(torch.isfinite(last["x"]["video"]).all().item() and torch.isfinite(last["x"]["audio"]).all().item())
    ms_total, ms_e2e = ctx.max_over_ranks(ms_total, ms_e2e)
    if rank != 0:
        return None
    K = runs * NFE
    value = world * B * K / (ms_total * 1e-3)
    res = _common_line(ctx, METRIC, value, K, NFE, ms_total, "DPM-Solver++ 50 NFE (predict_x0, thresholding, order 2, "
                       "time_uniform, multistep) — BASELINE.json configs[2]", B,
                       {"nfe_per_run": NFE, "runs_timed": runs, "parallelism": f"batch-shard x{world}"})
    res.update({"e2e": {"value": round(world * B * K / (ms_e2e * 1e-3), 3), "unit": UNIT, "ms_per_step": round(ms_e2e / K, 4),
                        "h2d_bytes_per_step": (xv_h.numel() + xa_h.numel()) * 4 // NFE,
                        "d2h_bytes_per_step": (xv_h.numel() + xa_h.numel()) * 4 // NFE,
                        "note": "x_T from pinned host memory and the finished sample back to the host once per 50-NFE run"},
                "gpu_launches": K * (model.num_launches(B) + 6),
                "roofline": _tensor_roof(FLOPS_PER_EVAL * B, ms_total / K, "whole NFE (model evaluation + solver update)"),
                "finite": finite})
    return res


def run_b200_cond(ctx, args, model, diffusion):
    """BASELINE.json configs[4]: audio -> video zero-shot conditional sampling by replacement (class_scale 0,
    multimodal_gaussian_diffusion.py:642-720), batch 4 per GPU (32 over 8 GPUs).  A step = overwrite the audio with
    q_sample(condition, t, fixed noise) + one ancestral p_sample step."""
    import torch
    world, rank, device = ctx.world, ctx.rank, ctx.device
    B, K, W = (args.batch if (args.batch_set and args.workload == "cond") else 4), args.steps, max(args.warmup, 3)
    import random
    random.seed(4321 + rank)
    gen = torch.Generator().manual_seed(1234 + rank)
    noise_h = {"video": torch.randn(B, *VIDEO_SIZE, generator=gen).pin_memory(),
               "audio": torch.randn(B, *AUDIO_SIZE, generator=gen).pin_memory()}
    cond_h = (0.1 * torch.randn(B, *AUDIO_SIZE, generator=gen)).pin_memory()
    shape = {"video": (B, *VIDEO_SIZE), "audio": (B, *AUDIO_SIZE)}
    noise = {k: v.to(device) for k, v in noise_h.items()}
    # device-resident: the library's own loop generator, W + K of its 1000 steps
    loop = diffusion.conditional_p_sample_loop_progressive_unscale(
        model, shape, use_fp16=True, noise=noise, model_kwargs={"audio": cond_h.to(device)}, device=device, class_scale=0.0)
    last = {}

    def dev_step(i):
        last["x"] = next(loop)
    ms_total = timed(ctx, dev_step, W, K)
    finite = bool(torch.isfinite(last["x"]["video"]).all().item())
    # end to end: the condition and the state come from pinned host memory every step, the sample goes back
    T = diffusion.num_timesteps
    hv_out, ha_out = torch.empty(B, *VIDEO_SIZE).pin_memory(), torch.empty(B, *AUDIO_SIZE).pin_memory()

    def e2e_step(i):
        t = torch.full((B,), (T - 1 - i) % T, device=device, dtype=torch.long)
        with torch.no_grad():
            c = cond_h.to(device, non_blocking=True)
            x = {"video": noise_h["video"].to(device, non_blocking=True),
                 "audio": diffusion.q_sample(c, t, noise=noise["audio"])}
            s = diffusion.p_sample(model, x, t)["sample"]
        hv_out.copy_(s["video"], non_blocking=True)
        ha_out.copy_(s["audio"], non_blocking=True)
    ms_e2e = timed(ctx, e2e_step, W, K)
    ms_total, ms_e2e = ctx.max_over_ranks(ms_total, ms_e2e)
    if rank != 0:
        return None
    value = world * B * K / (ms_total * 1e-3)
    res = _common_line(ctx, METRIC, value, K, W, ms_total, "audio->video replacement-conditioned p_sample step "
                       "(conditional_p_sample_loop, class_scale 0) — BASELINE.json configs[4]", B,
                       {"parallelism": f"batch-shard x{world} (global batch {B * world})"})
    nb = (noise_h["video"].numel() + cond_h.numel()) * 4
    res.update({"e2e": {"value": round(world * B * K / (ms_e2e * 1e-3), 3), "unit": UNIT, "ms_per_step": round(ms_e2e / K, 4),
                        "h2d_bytes_per_step": nb, "d2h_bytes_per_step": (noise_h["video"].numel() + noise_h["audio"].numel()) * 4},
                "gpu_launches": K * (model.num_launches(B) + 5),
                "roofline": _tensor_roof(FLOPS_PER_EVAL * B, ms_total / K, "whole step (q_sample overwrite + model evaluation + tail)"),
                "finite": finite})
    return res


# ----------------------------------------------------------------------------- CPU arms (oracle = checker / baseline only)
def usable_cores():
    """Host cores this process may actually use: affinity mask, capped by the cgroup CPU quota if there is one."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:
            quota, period = f.read().split()
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except Exception:
        pass
    return max(1, n)


def _oracle_setup():
    import torch
    from oracle.mmdiff_oracle import DiffusionOracle, UNetConfig, draw_shifts, synthetic_state_dict
    torch.set_num_threads(min(usable_cores(), 64))  # MKL-DNN stops scaling (and oversubscribes) beyond that
    cfg = UNetConfig()
    sd = synthetic_state_dict(cfg, seed=0)
    return cfg, sd, DiffusionOracle(1000), draw_shifts


def _oracle_step(cfg, sd, diff, draw_shifts, batch, seed):
    import random
    import torch
    g = torch.Generator().manual_seed(seed)
    x = {"video": torch.randn(batch, *VIDEO_SIZE, generator=g), "audio": torch.randn(batch, *AUDIO_SIZE, generator=g)}
    z = {"video": torch.randn(batch, *VIDEO_SIZE, generator=g), "audio": torch.randn(batch, *AUDIO_SIZE, generator=g)}
    t = torch.full((batch,), 500, dtype=torch.long)
    t0 = time.perf_counter()
    with torch.no_grad():
        diff.p_sample(sd, cfg, x, t, z, draw_shifts(cfg, random.Random(seed)))
    return time.perf_counter() - t0


def _reference_cpu_setup():
    """(step_fn(batch, seed) -> seconds, kind, threads): one p_sample step of the reference's own CPU implementation —
    the UNMODIFIED reference from baseline/_ref when present ("reference"), else the oracle port ("port")."""
    import random
    import torch
    threads = min(usable_cores(), 64)   # MKL-DNN stops scaling (and oversubscribes) beyond that
    torch.set_num_threads(threads)
    su = import_reference()
    if su is not None:
        model, diffusion = build_reference(su, torch.device("cpu"), use_fp16=False)

        def step(batch, seed):
            g = torch.Generator().manual_seed(seed)
            x = {"video": torch.randn(batch, *VIDEO_SIZE, generator=g), "audio": torch.randn(batch, *AUDIO_SIZE, generator=g)}
            t = torch.full((batch,), 500, dtype=torch.long)
            random.seed(seed)
            t0 = time.perf_counter()
            with torch.no_grad():
                diffusion.p_sample(model, x, t)
            return time.perf_counter() - t0
        return step, "reference", threads
    cfg, sd, diff, draw = _oracle_setup()
    return (lambda batch, seed: _oracle_step(cfg, sd, diff, draw, batch, seed)), "port", threads


def cpu_baseline(batch, max_seconds=45.0):
    """The reference's CPU path timed on the host cores on a bounded sample of the SAME workload (batch `batch` p_sample
    steps, fp32): 1 warm-up + up to 2 timed steps inside the time budget."""
    step, kind, threads = _reference_cpu_setup()
    start = time.perf_counter()
    times = [step(batch, 0)]  # doubles as warm-up; replaced if there is time for more
    for i in range(2):
        if time.perf_counter() - start + times[-1] > max_seconds:
            break
        t = step(batch, i + 1)
        times = [t] if i == 0 else times + [t]
    med = statistics.median(times)
    return {"value": round(batch / med, 4), "unit": UNIT, "cores": threads, "kind": kind,
            "sample": f"{len(times)} timed p_sample step(s) at batch {batch} (the bench workload) after "
                      f"{'1 warm-up' if len(times) > 1 or time.perf_counter() - start > times[0] * 1.5 else 'no warm-up'}, fp32, "
                      f"median {med:.2f} s/step"}


def run_reference(args):
    """bench.py --impl reference: the reference's own CPU implementation of the path (host cores, all usable threads) on
    this arm's config / metric / unit; each step one p_sample step at the bench batch, bounded by a time budget."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    B, K, W = args.batch, args.steps, max(args.warmup, 1)
    step, kind, threads = _reference_cpu_setup()
    budget = 240.0
    start = time.perf_counter()
    first = step(B, 0)  # warm-up (also sizes the run)
    n_warm = 1
    if W > 1 and first * (K + 2) < budget:
        step(B, 1)
        n_warm = 2
    times = []
    for i in range(K):
        times.append(step(B, 100 + i))
        if time.perf_counter() - start + times[-1] > budget:
            break
    total = sum(times)
    value = B * len(times) / total
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": world,
            "steps": len(times), "steps_requested": K, "warmup": n_warm, "ms_per_step": round(1e3 * total / len(times), 2),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "p_sample step of the 1000-step DDPM loop (BASELINE.json configs[1])",
                       "batch_per_gpu": B, "global_batch": B, "video": VIDEO_SIZE, "audio": AUDIO_SIZE,
                       "sample": f"whole batch-{B} p_sample steps on the host cores, bounded by a {budget:.0f} s budget"},
            "cpu_baseline": {"value": round(value, 4), "unit": UNIT, "cores": threads, "kind": kind,
                             "sample": f"{len(times)} p_sample steps at batch {B} on the host cores (time budget {budget:.0f} s)"},
            "e2e": {"value": round(value, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=None, help="samples per GPU (sample / cond: 4, dpm: 16, train: 8)")
    ap.add_argument("--workload", default="sample", choices=["sample", "dpm", "cond", "train"],
                    help="sample: p_sample step of configs[1] (the headline metric, with the others as `extra` sub-records); "
                         "dpm / cond / train: configs[2] / configs[4] / configs[3] alone")
    ap.add_argument("--profile-reps", type=int, default=3)
    ap.add_argument("--train-steps", type=int, default=0, help="timed steps of the train sub-record (default: 8 as an extra, --steps alone)")
    ap.add_argument("--dropout", type=float, default=0.1, help="train workload: dropout of the shipped training flags (ssh_scripts/multimodal_train.sh)")
    ap.add_argument("--ddp", action="store_true", help="train workload, N>1: torch DistributedDataParallel instead of the flat all-reduce")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the reference-through-PyTorch-eager GPU baseline")
    ap.add_argument("--no-extras", action="store_true", help="skip the dpm / cond / train sub-records of the default line")
    args = ap.parse_args()
    args.batch_set = args.batch is not None
    if args.batch is None:
        args.batch = 4
    if args.impl == "reference":
        run_reference(args)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus != world and world == 1 and args.gpus > 1:
        raise SystemExit(f"--gpus {args.gpus} needs torchrun (one process per GPU); WORLD_SIZE is {world}")
    ctx = Ctx()
    model, diffusion = build_b200(ctx.device)
    if args.workload == "sample":
        result = run_b200(ctx, args, model, diffusion)
        if not args.no_extras:
            extra = {}
            for name, fn in (("dpm", run_b200_dpm), ("cond", run_b200_cond), ("train", run_b200_train)):
                sub_args = argparse.Namespace(**vars(args))
                if name == "train":
                    sub_args.train_steps = args.train_steps or 8
                    sub_args.profile_reps = 0
                try:
                    extra[name] = fn(ctx, sub_args, model, diffusion)
                except Exception as e:  # noqa: BLE001  (a sub-record must not take the headline line down)
                    extra[name] = {"error": f"{type(e).__name__}: {str(e)[:300]}"}
            if result is not None:
                result["extra"] = extra
    else:
        fn = {"dpm": run_b200_dpm, "cond": run_b200_cond, "train": run_b200_train}[args.workload]
        result = fn(ctx, args, model, diffusion)
    ctx.close()
    if ctx.rank == 0:
        print(json.dumps(result))


if __name__ == "__main__":
    main()
