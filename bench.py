#!/usr/bin/env python
"""Benchmark of the MM-Diffusion denoising hot path on B200 (contract: see task prompt / DESIGN.md §measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A "step" = one p_sample step (MultimodalUNet.forward + sampler tail) over one batch of synthetic input.
Workload at N=1 = BASELINE.json configs[1]: the per-step unit of the 1000-step DDPM p_sample_loop at batch 4,
16x3x64x64 video + 1x25600 audio, random-init production U-Net (133.7 M params).  N>1: one process per GPU
(torchrun), batch 4 per rank (weak scaling), no collective inside a step; the finished samples are gathered
once with NCCL at the end of the timed region.  Metric: denoising steps/sec = samples x steps / time.

--impl reference times the reference's CPU algorithm (the oracle port, oracle/mmdiff_oracle.py — the unmodified
Python reference cannot travel to the GPU box) on the host cores for the same metric.
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

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
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
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


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
    path = os.path.join(ROOT, "profiles", "r01_dram_traffic.json")
    try:
        with open(path) as f:
            t = json.load(f)
        if t.get("batch") != batch:
            return None
        ks = t["kernels"]
        return (ks.get(kernel) or ks["mmd::" + kernel])["dram_bytes_per_launch"]
    except Exception:
        return None


def family_summary(steps):
    fam = {}
    for s in steps:
        f = fam.setdefault(s["kind"], {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
        f["ms"] += s["ms"]; f["flops"] += s["flops"]; f["bytes"] += s["bytes"]; f["launches"] += s["kernels"]
    return fam


def run_b200(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    B, K, W = args.batch, args.steps, max(args.warmup, 3)
    model, diffusion = build_b200(device)
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
        v8 = ((x["video"] + 1) * 127.5).clamp(0, 255).to(torch.uint8)
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
        for i in range(W):
            x = diffusion.p_sample(model, x, ts[i])["sample"]
        sample_epilogue(x)   # warm the end-of-loop path too (first NCCL call builds the communicator)
        barrier()
        clocks = ClockSampler(local_rank)
        if rank == 0:
            clocks.start()
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
        traffic = ncu_traffic(dname, B)
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
                                     "summarised in profiles/r01_dram_traffic.json" if traffic is not None else None,
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
        if world == 1 and not args.no_cpu_baseline:
            result["cpu_baseline"] = cpu_baseline(max_seconds=40.0)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(result))


def run_b200_train(args):
    """BASELINE.json configs[3]: multimodal_training_losses forward + backward, batch 8 per GPU, synthetic
    Landscape-shape data, batch-sharded over the GPUs of the box (DDP gradient all-reduce over NCCL when N > 1).
    A step = one forward + backward over one batch (no optimizer: the config names fwd+bwd); value = samples x steps / s."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    B, K, W = (args.batch if args.batch_set else 8), args.steps, max(args.warmup, 3)
    model, diffusion = build_b200(device)
    model.convert_to_fp32()
    model.train()
    net = model
    if world > 1 and args.ddp:
        from torch.nn.parallel import DistributedDataParallel as DDP
        net = DDP(model, device_ids=[local_rank], broadcast_buffers=False, bucket_cap_mb=128)
    from mm_diffusion_b200.parallel import allreduce_flat_gradients
    import random
    random.seed(4321 + rank)
    gen = torch.Generator().manual_seed(1234 + rank)
    xv_h = torch.randn(B, *VIDEO_SIZE, generator=gen).clamp(-1, 1).pin_memory()
    xa_h = torch.randn(B, *AUDIO_SIZE, generator=gen).clamp(-1, 1).pin_memory()
    t = torch.randint(0, diffusion.num_timesteps, (B,), generator=torch.Generator().manual_seed(99 + rank)).to(device)
    torch.manual_seed(7 + rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(x0):
        model.zero_grad(set_to_none=True)
        terms = diffusion.multimodal_training_losses(net, x0, t)
        loss = terms["loss"].mean()
        loss.backward()
        if world > 1 and not args.ddp:
            allreduce_flat_gradients(model)   # the step's one collective: flat fp32 gradient buffer over NCCL
        return loss

    x0 = {"video": xv_h.to(device), "audio": xa_h.to(device)}
    for _ in range(W):
        step(x0)
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        loss = step(x0)
    e1.record()
    barrier()
    clk = clocks.stop() if rank == 0 else None
    ms_total = e0.elapsed_time(e1)
    finite = bool(torch.isfinite(loss).item() and torch.isfinite(model.flat_grad).all().item())
    # end to end: host batch -> device, step, loss back to the host
    for _ in range(W):
        step({"video": xv_h.to(device, non_blocking=True), "audio": xa_h.to(device, non_blocking=True)}).item()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(K):
        step({"video": xv_h.to(device, non_blocking=True), "audio": xa_h.to(device, non_blocking=True)}).item()
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    if world > 1:
        tmax = torch.tensor([ms_total, ms_e2e], device=device)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms_total, ms_e2e = tmax[0].item(), tmax[1].item()
    if rank == 0:
        peaks = load_peaks()
        fwd_steps = model.plan_steps(B)   # inference plan may be absent; launch counts come from the training plan below
        bwd = model.profile_backward(B, reps=1) if args.profile_reps > 0 else []
        fam = {}
        for s_ in bwd:
            f = fam.setdefault(s_["kind"], {"ms": 0.0, "steps": 0})
            f["ms"] += s_["ms"]; f["steps"] += 1
        flops_fwd = 1.3288e12 * B   # algorithmic FLOPs per model evaluation per sample
        step_flops = 3.0 * flops_fwd  # forward + dgrad + wgrad
        value = world * B * K / (ms_total * 1e-3)
        result = {
            "metric": "training_losses forward+backward samples/sec (16fx64x64 video + 25600 audio)", "value": round(value, 3),
            "unit": "sample-steps/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": round(ms_total / K, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": "multimodal_training_losses forward+backward (BASELINE.json configs[3])",
                       "batch_per_gpu": B, "global_batch": B * world, "video": VIDEO_SIZE, "audio": AUDIO_SIZE,
                       "params_m": round(sum(p.numel() for p in model.parameters()) / 1e6, 2),
                       "parallelism": f"batch-shard x{world}" + ((" (torch DDP buckets over NCCL)" if args.ddp else
                                                                  " (one flat fp32 gradient all-reduce over NCCL per step)") if world > 1 else ""),
                       "l2_note": "kept activations + gradients (~48 GB at B=8) exceed the 126 MB L2",
                       "cuda_graph": os.environ.get("MMD_TRAIN_GRAPH", "1") != "0"},
            "e2e": {"value": round(world * B * K / (ms_e2e * 1e-3), 3), "unit": "sample-steps/s", "ms_per_step": round(ms_e2e / K, 3),
                    "h2d_bytes_per_step": (xv_h.numel() + xa_h.numel()) * 4, "d2h_bytes_per_step": 4},
            "gpu_launches": K * (model.num_backward_launches(B) + len(fwd_steps)),
            "model_tflops": round(step_flops * K / (ms_total * 1e-3) / 1e12, 2),
            "roofline": {"bound": "tensor", "achieved": round(step_flops * K / (ms_total * 1e-3) / 1e12, 2),
                         "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                         "frac": round(step_flops * K / (ms_total * 1e-3) / 1e12 / peaks["bf16_tflops_sustained"], 4),
                         "kernel": "whole step (forward + dgrad + wgrad, 3 x 1.3288 TFLOP per sample)", "traffic": None},
            "backward_families": {k: {"ms": round(v["ms"], 3), "steps": v["steps"]} for k, v in
                                  sorted(fam.items(), key=lambda kv: -kv[1]["ms"])},
            "clocks": clk, "finite": finite,
        }
        print(json.dumps(result))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


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


def cpu_baseline(max_seconds=40.0):
    """The reference algorithm (oracle port, PyTorch fp32 on MKL-DNN) timed on the host cores on a bounded sample:
    single-sample p_sample steps (1 warm-up + up to 3 timed, stops at the time budget)."""
    import torch
    cfg, sd, diff, draw = _oracle_setup()
    start = time.perf_counter()
    times = [_oracle_step(cfg, sd, diff, draw, 1, 0)]  # doubles as warm-up; replaced if there is time for more
    for i in range(3):
        if time.perf_counter() - start > max_seconds:
            break
        t = _oracle_step(cfg, sd, diff, draw, 1, i + 1)
        times = [t] if i == 0 else times + [t]
    med = statistics.median(times)
    return {"value": round(1.0 / med, 4), "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{len(times)} timed single-sample p_sample steps (batch 1 of the batch-4 workload), fp32, "
                      f"median {med:.2f} s/step"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    K, W = args.steps, max(args.warmup, 1)
    cfg, sd, diff, draw = _oracle_setup()
    import torch
    budget = 200.0
    start = time.perf_counter()
    first = _oracle_step(cfg, sd, diff, draw, 1, 0)  # warm-up (also sizes the run)
    n_warm = 1
    if W > 1 and first < 20.0:
        _oracle_step(cfg, sd, diff, draw, 1, 1)
        n_warm = 2
    times = []
    for i in range(K):
        times.append(_oracle_step(cfg, sd, diff, draw, 1, 100 + i))
        if time.perf_counter() - start > budget:
            break
    total = sum(times)
    value = len(times) / total
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": world,
            "steps": len(times), "steps_requested": K, "warmup": n_warm, "ms_per_step": round(1e3 * total / len(times), 2),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "p_sample step of the 1000-step DDPM loop (BASELINE.json configs[1])",
                       "video": VIDEO_SIZE, "audio": AUDIO_SIZE, "sample": "one sample of the batch per step (bounded)"},
            "cpu_baseline": {"value": round(value, 4), "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"{len(times)} single-sample p_sample steps on the host cores (time budget {budget:.0f} s)"},
            "e2e": {"value": round(value, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=None, help="samples per GPU (configs[1] uses 4, the training config 8)")
    ap.add_argument("--workload", default="sample", choices=["sample", "train"],
                    help="sample: p_sample step of configs[1] (the headline metric); train: training_losses fwd+bwd of configs[3]")
    ap.add_argument("--profile-reps", type=int, default=3)
    ap.add_argument("--ddp", action="store_true", help="train workload, N>1: torch DistributedDataParallel instead of the flat all-reduce")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.batch_set = args.batch is not None
    if args.batch is None:
        args.batch = 4
    if args.impl == "reference":
        run_reference(args)
    else:
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if args.gpus != world and world == 1 and args.gpus > 1:
            raise SystemExit(f"--gpus {args.gpus} needs torchrun (one process per GPU); WORLD_SIZE is {world}")
        if args.workload == "train":
            run_b200_train(args)
        else:
            run_b200(args)


if __name__ == "__main__":
    main()
