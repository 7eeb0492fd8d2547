"""CPU oracle for the MM-Diffusion denoising hot path.  TEST INFRASTRUCTURE ONLY.

A plain PyTorch fp32 restatement of the reference algorithm, written from the
reference's behaviour (file:line citations are relative to the reference repo
researchmm/MM-Diffusion).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; the product path
(mm_diffusion_b200/) never does.

Parity pinning: the reference ships no golden vectors or tests, so this
oracle is pinned against the reference *itself*, imported unmodified in the build
container by oracle/make_golden.py, which writes tests/golden/*.pt; the CPU test
tests/test_oracle_golden.py replays those fixtures through this file.

Functional style: every function takes the reference's state_dict (same key names,
SURVEY.md App. F) instead of owning parameters.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- config / topology
@dataclass
class UNetConfig:
    """Constructor arguments of MultimodalUNet (multimodal_unet.py:737-764) after create_model's
    string parsing (multimodal_script_util.py:156-201)."""
    video_size: Sequence[int] = (16, 3, 64, 64)
    audio_size: Sequence[int] = (1, 25600)
    model_channels: int = 128
    video_out_channels: int = 3
    audio_out_channels: int = 1
    num_res_blocks: int = 2
    channel_mult: Sequence[int] = (1, 2, 3, 4)
    num_heads: int = 4
    num_head_channels: int = 64
    cross_attention_resolutions: Sequence[int] = (2, 4, 8)
    cross_attention_windows: Sequence[int] = (1, 4, 8)
    cross_attention_shift: bool = True
    video_attention_resolutions: Sequence[int] = (2, 4, 8)
    audio_attention_resolutions: Sequence[int] = (-1,)


@dataclass
class ResSpec:
    prefix: str
    cin: int
    cout: int
    dilation: int
    up: bool = False
    down: bool = False
    video_attention: bool = False
    audio_attention: bool = False


@dataclass
class CrossSpec:
    prefix: str
    channels: int
    heads: int
    window: int
    shift: bool


@dataclass
class Topology:
    input_blocks: List[list] = field(default_factory=list)   # list of lists of specs (block 0 = "initial")
    middle: list = field(default_factory=list)
    output_blocks: List[list] = field(default_factory=list)
    final_ch: int = 0

    def cross_specs(self) -> List[CrossSpec]:
        out = []
        for blk in self.input_blocks + [self.middle] + self.output_blocks:
            out += [s for s in blk if isinstance(s, CrossSpec)]
        return out


def build_topology(cfg: UNetConfig) -> Topology:
    """Restates the constructor's block schedule (multimodal_unet.py:799-1012): channel plan,
    audio dilation counter 2**(i % 10), where attention / cross-attention / up / down blocks sit."""
    topo = Topology()
    mc = cfg.model_channels
    ch = int(cfg.channel_mult[0] * mc)
    chans = [ch]
    topo.input_blocks.append(["initial"])
    ds, dil = 1, 1
    cross_heads = lambda c: cfg.num_heads if cfg.num_head_channels == -1 else c // cfg.num_head_channels

    def cross(prefix, c, ds_):
        i = list(cfg.cross_attention_resolutions).index(ds_)
        return CrossSpec(prefix, c, cross_heads(c), cfg.cross_attention_windows[i], bool(cfg.cross_attention_shift))

    for level, mult in enumerate(cfg.channel_mult):
        for _ in range(cfg.num_res_blocks):
            idx = len(topo.input_blocks)
            cout = int(mult * mc)
            blk = [ResSpec(f"input_blocks.{idx}.0", ch, cout, 2 ** (dil % 10),
                           video_attention=ds in cfg.video_attention_resolutions,
                           audio_attention=ds in cfg.audio_attention_resolutions)]
            dil += 1
            ch = cout
            if ds in cfg.cross_attention_resolutions:
                blk.append(cross(f"input_blocks.{idx}.1", ch, ds))
            topo.input_blocks.append(blk)
            chans.append(ch)
        if level != len(cfg.channel_mult) - 1:
            idx = len(topo.input_blocks)
            topo.input_blocks.append([ResSpec(f"input_blocks.{idx}.0", ch, ch, 2 ** (dil % 10), down=True)])
            dil += 1
            chans.append(ch)
            ds *= 2
    mid_dil = 2 ** (dil % 10)
    if list(cfg.cross_attention_windows) == [1, 4, 8]:  # multimodal_unet.py:875
        topo.middle = [
            ResSpec("middle_blocks.0", ch, ch, mid_dil, video_attention=True, audio_attention=True),
            CrossSpec("middle_blocks.1", ch, cross_heads(ch), cfg.video_size[0], False),
            ResSpec("middle_blocks.2", ch, ch, mid_dil, video_attention=True, audio_attention=True),
        ]
    else:
        topo.middle = [
            ResSpec("middle_blocks.0", ch, ch, mid_dil, video_attention=True, audio_attention=True),
            ResSpec("middle_blocks.1", ch, ch, mid_dil, video_attention=True, audio_attention=True),
        ]
    dil -= 1
    for level, mult in list(enumerate(cfg.channel_mult))[::-1]:
        for i in range(cfg.num_res_blocks + 1):
            idx = len(topo.output_blocks)
            ich = chans.pop()
            cout = int(mc * mult)
            blk = [ResSpec(f"output_blocks.{idx}.0", ch + ich, cout, 2 ** (dil % 10),
                           video_attention=ds in cfg.video_attention_resolutions,
                           audio_attention=ds in cfg.audio_attention_resolutions)]
            dil -= 1
            ch = cout
            if ds in cfg.cross_attention_resolutions:
                blk.append(cross(f"output_blocks.{idx}.{len(blk)}", ch, ds))
            if level and i == cfg.num_res_blocks:
                blk.append(ResSpec(f"output_blocks.{idx}.{len(blk)}", ch, ch, 2 ** (dil % 10), up=True))
                ds //= 2
            topo.output_blocks.append(blk)
    topo.final_ch = ch
    return topo


# --------------------------------------------------------------------------- primitives
def timestep_embedding(t: torch.Tensor, dim: int, max_period: float = 10000.0) -> torch.Tensor:
    """nn.py:192-210: [cos(t f_i), sin(t f_i)], f_i = max_period^(-i/half)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def group_norm32(x: torch.Tensor, sd, prefix: str, channel_dim: int = 1) -> torch.Tensor:
    """nn.py:16-33: GroupNorm(32, C), eps 1e-5, statistics over everything but (batch, group).
    `channel_dim` says where C sits; video tensors are [B,F,C,H,W] (channel_dim=2)."""
    w, b = sd[prefix + ".GroupNorm.weight"], sd[prefix + ".GroupNorm.bias"]
    if channel_dim != 1:
        x = x.movedim(channel_dim, 1)
    y = F.group_norm(x, 32, w, b, eps=1e-5)
    if channel_dim != 1:
        y = y.movedim(1, channel_dim)
    return y


def video_conv_2d1d(v: torch.Tensor, sd, prefix: str) -> torch.Tensor:
    """VideoConv '2d+1d' (multimodal_unet.py:91-99): per-frame 3x3 conv, then k=3 conv
    along frames per pixel, both 'same' zero padding, both biased.  v: [B,F,C,H,W]."""
    B, Fr, C, H, W = v.shape
    u = F.conv2d(v.reshape(B * Fr, C, H, W), sd[prefix + ".video_conv_spatial.weight"],
                 sd[prefix + ".video_conv_spatial.bias"], padding=1)
    Co = u.shape[1]
    u = u.reshape(B, Fr, Co, H, W).permute(0, 3, 4, 2, 1).reshape(B * H * W, Co, Fr)
    y = F.conv1d(u, sd[prefix + ".video_conv_temporal.weight"], sd[prefix + ".video_conv_temporal.bias"], padding=1)
    return y.reshape(B, H, W, Co, Fr).permute(0, 4, 3, 1, 2)


def video_conv_3d(v: torch.Tensor, sd, prefix: str) -> torch.Tensor:
    """VideoConv '3d' (multimodal_unet.py:101-104): Conv3d on [B,C,F,H,W] with 'same' padding."""
    w = sd[prefix + ".video_conv.weight"]
    pad = tuple(k // 2 for k in w.shape[2:])
    return F.conv3d(v.permute(0, 2, 1, 3, 4), w, sd[prefix + ".video_conv.bias"], padding=pad).permute(0, 2, 1, 3, 4)


def audio_conv(a: torch.Tensor, sd, prefix: str, dilation: int = 1) -> torch.Tensor:
    """AudioConv (multimodal_unet.py:108-131): Conv1d, 'same' padding = dilation*(k-1)/2."""
    w = sd[prefix + ".audio_conv.weight"]
    k = w.shape[-1]
    return F.conv1d(a, w, sd[prefix + ".audio_conv.bias"], padding=dilation * (k - 1) // 2, dilation=dilation)


def qkv_attention(qkv: torch.Tensor, heads: int) -> torch.Tensor:
    """SingleModalQKVAttention.forward (multimodal_unet.py:221-240): qkv [N, 3*H*d, T]; channels
    split q|k|v, head h = channels [h*d,(h+1)*d); softmax(q^T k / sqrt(d)) in fp32."""
    N, W3, T = qkv.shape
    d = W3 // (3 * heads)
    q, k, v = qkv.chunk(3, dim=1)
    q = q.reshape(N * heads, d, T)
    k = k.reshape(N * heads, d, T)
    v = v.reshape(N * heads, d, T)
    w = torch.einsum("bct,bcs->bts", q, k) / math.sqrt(d)
    w = torch.softmax(w.float(), dim=-1)
    return torch.einsum("bts,bcs->bct", w, v).reshape(N, heads * d, T)


def single_modal_attention(x: torch.Tensor, sd, prefix: str, heads: int) -> torch.Tensor:
    """SingleModalAtten._forward (multimodal_unet.py:280-287): x [N,C,T] -> x + proj(attn(qkv(norm(x))))."""
    h = group_norm32(x, sd, prefix + ".norm")
    qkv = F.conv1d(h, sd[prefix + ".qkv.weight"], sd[prefix + ".qkv.bias"])
    h = qkv_attention(qkv, heads)
    h = F.conv1d(h, sd[prefix + ".proj_out.weight"], sd[prefix + ".proj_out.bias"])
    return x + h


def res_block(v, a, emb, sd, s: ResSpec, heads: int, drop=None):
    """ResBlock._forward (multimodal_unet.py:434-495), use_scale_shift_norm=True.
    v [B,F,C,H,W], a [B,C,L], emb [B,E].  drop = None (eval: Dropout is the identity) or (video keep mask, audio keep
    mask, scale): nn.Dropout between the SiLU and the out conv (:376,384) with the masks supplied by the caller."""
    p = s.prefix
    B, Fr, C, H, W = v.shape
    vh = video_conv_2d1d(F.silu(group_norm32(v, sd, p + ".video_in_layers.0", 2)), sd, p + ".video_in_layers.2")
    ah = audio_conv(F.silu(group_norm32(a, sd, p + ".audio_in_layers.0")), sd, p + ".audio_in_layers.2", s.dilation)
    if s.down:      # conv at source resolution, then pool both branches (:441-448)
        pool_v = lambda t: F.avg_pool2d(t.reshape(-1, *t.shape[2:]), 2).reshape(B, Fr, t.shape[2], H // 2, W // 2)
        vh, v = pool_v(vh), pool_v(v)
        ah, a = F.avg_pool1d(ah, 4), F.avg_pool1d(a, 4)
    elif s.up:
        up_v = lambda t: F.interpolate(t.reshape(-1, *t.shape[2:]), scale_factor=2, mode="nearest").reshape(
            B, Fr, t.shape[2], H * 2, W * 2)
        vh, v = up_v(vh), up_v(v)
        ah, a = F.interpolate(ah, scale_factor=4, mode="nearest"), F.interpolate(a, scale_factor=4, mode="nearest")
    e = F.linear(F.silu(emb), sd[p + ".emb_layers.1.weight"], sd[p + ".emb_layers.1.bias"])
    scale, shift = e.chunk(2, dim=1)
    vh = group_norm32(vh, sd, p + ".video_out_layers.0", 2) * (1 + scale[:, None, :, None, None]) + shift[:, None, :, None, None]
    vh = F.silu(vh)
    if drop is not None:
        vh = vh * drop[0].to(vh.dtype) * drop[2]
    vh = video_conv_3d(vh, sd, p + ".video_out_layers.3")
    ah = group_norm32(ah, sd, p + ".audio_out_layers.0") * (1 + scale[:, :, None]) + shift[:, :, None]
    ah = F.silu(ah)
    if drop is not None:
        ah = ah * drop[1].to(ah.dtype) * drop[2]
    ah = audio_conv(ah, sd, p + ".audio_out_layers.3")
    if s.cin != s.cout:
        v = video_conv_3d(v, sd, p + ".video_skip_connection")
        a = audio_conv(a, sd, p + ".audio_skip_connection")
    v = v + vh
    a = a + ah
    if s.video_attention:  # spatial per frame, then temporal per pixel (:485-491)
        B, Fr, C, H, W = v.shape
        x = v.permute(0, 1, 2, 3, 4).reshape(B * Fr, C, H * W)
        x = single_modal_attention(x, sd, p + ".spatial_attention_block", heads)
        x = x.reshape(B, Fr, C, H, W).permute(0, 3, 4, 2, 1).reshape(B * H * W, C, Fr)
        x = single_modal_attention(x, sd, p + ".temporal_attention_block", heads)
        v = x.reshape(B, H, W, C, Fr).permute(0, 4, 3, 1, 2)
    if s.audio_attention:
        a = single_modal_attention(a, sd, p + ".audio_attention_block", heads)
    return v, a


def cross_attention(v, a, sd, s: CrossSpec, shift: int):
    """CrossAttentionBlock._forward + QKVAttention.forward (multimodal_unet.py:507-564, 614-678;
    SURVEY.md C-4).  Video tokens of frame i attend the audio tokens of segments (i+shift+j) mod F,
    j < window; audio tokens of segment i attend the video tokens of frames (i+shift+j) mod F.
    K/V of video queries come from the audio projection and vice versa."""
    p = s.prefix
    B, Fr, C, H, W = v.shape
    L = a.shape[2]
    hw, apf = H * W, L // Fr
    assert apf * Fr == L, "oracle restates the exact-division case only (production shapes)"
    heads, d = s.heads, C // s.heads
    vt = v.permute(0, 2, 1, 3, 4).reshape(B, C, Fr * hw)
    vq = F.conv1d(group_norm32(vt, sd, p + ".v_norm"), sd[p + ".v_qkv.weight"], sd[p + ".v_qkv.bias"])
    aq = F.conv1d(group_norm32(a, sd, p + ".a_norm"), sd[p + ".a_qkv.weight"], sd[p + ".a_qkv.bias"])
    vQ, vK, vV = [t.reshape(B, heads, d, Fr, hw) for t in vq.chunk(3, dim=1)]
    aQ, aK, aV = [t.reshape(B, heads, d, Fr, apf) for t in aq.chunk(3, dim=1)]
    v_out = torch.empty_like(vQ)
    a_out = torch.empty_like(aQ)
    scale = 1.0 / math.sqrt(d)
    for i in range(Fr):
        blocks = [(i + shift + j) % Fr for j in range(s.window)]
        k = aK[:, :, :, blocks].reshape(B, heads, d, -1)
        val = aV[:, :, :, blocks].reshape(B, heads, d, -1)
        w = torch.softmax(torch.einsum("bhdq,bhdk->bhqk", vQ[:, :, :, i], k) * scale, dim=-1)
        v_out[:, :, :, i] = torch.einsum("bhqk,bhdk->bhdq", w, val)
        k = vK[:, :, :, blocks].reshape(B, heads, d, -1)
        val = vV[:, :, :, blocks].reshape(B, heads, d, -1)
        w = torch.softmax(torch.einsum("bhdq,bhdk->bhqk", aQ[:, :, :, i], k) * scale, dim=-1)
        a_out[:, :, :, i] = torch.einsum("bhqk,bhdk->bhdq", w, val)
    vh = v_out.reshape(B, C, Fr, H, W).permute(0, 2, 1, 3, 4)
    vh = video_conv_3d(vh, sd, p + ".video_proj_out")
    ah = audio_conv(a_out.reshape(B, C, L), sd, p + ".audio_proj_out")
    return v + vh, a + ah


def unet_forward(sd: Dict[str, torch.Tensor], cfg: UNetConfig, video, audio, t, shifts: Sequence[int], dropout=None):
    """MultimodalUNet.forward (multimodal_unet.py:1058-1101).  `shifts`: one window shift per
    CrossAttentionBlock in execution order (0 where the block does not shift), i.e. the values
    random.randint(0, F - window) returns inside attention_index (:619-622).
    dropout (training mode): {"scale": 1 / (1 - p), "masks": [(video keep mask, audio keep mask), ...]} with one pair
    per ResBlock in execution order."""
    topo = build_topology(cfg)
    shifts = list(shifts)
    masks = list(dropout["masks"]) if dropout else None
    emb = timestep_embedding(t, cfg.model_channels)
    emb = F.linear(F.silu(F.linear(emb, sd["time_embed.0.weight"], sd["time_embed.0.bias"])),
                   sd["time_embed.2.weight"], sd["time_embed.2.bias"])
    v, a = video.float(), audio.float()
    vs, as_ = [], []

    def run(blk, v, a):
        for s in blk:
            if s == "initial":
                v = video_conv_2d1d(v, sd, "input_blocks.0.0.video_conv")
                a = audio_conv(a, sd, "input_blocks.0.0.audio_conv")
            elif isinstance(s, ResSpec):
                drop = None
                if masks is not None:
                    mv, ma = masks.pop(0)
                    drop = (mv, ma, float(dropout["scale"]))
                v, a = res_block(v, a, emb, sd, s, cfg.num_heads, drop)
            else:
                v, a = cross_attention(v, a, sd, s, shifts.pop(0))
        return v, a

    for blk in topo.input_blocks:
        v, a = run(blk, v, a)
        vs.append(v)
        as_.append(a)
    v, a = run(topo.middle, v, a)
    for blk in topo.output_blocks:
        v = torch.cat([v, vs.pop()], dim=2)
        a = torch.cat([a, as_.pop()], dim=1)
        v, a = run(blk, v, a)
    v = video_conv_3d(F.silu(group_norm32(v, sd, "video_out.0", 2)), sd, "video_out.2")
    a = audio_conv(F.silu(group_norm32(a, sd, "audio_out.0")), sd, "audio_out.2")
    assert not shifts and not masks
    return v, a


def shift_bounds(cfg: UNetConfig) -> List[int]:
    """Upper bound (inclusive) of the randint draw of each cross-attention block in execution order;
    blocks with window_shift=False draw nothing (bound -1)."""
    out = []
    for s in build_topology(cfg).cross_specs():
        out.append(cfg.video_size[0] - s.window if s.shift else -1)
    return out


def draw_shifts(cfg: UNetConfig, rng) -> List[int]:
    """Draw shifts exactly as one reference forward would from Python's `random` API (rng = random
    module or random.Random): one randint per shifting block, in execution order."""
    return [rng.randint(0, b) if b >= 0 else 0 for b in shift_bounds(cfg)]


# --------------------------------------------------------------------------- diffusion math
class DiffusionOracle:
    """Tables and steps of GaussianDiffusion (multimodal_gaussian_diffusion.py:117-168) for the production
    setting: linear betas, EPSILON prediction, FIXED_LARGE variance, MSE loss, no respacing."""

    def __init__(self, steps: int = 1000):
        scale = 1000.0 / steps
        base = np.linspace(scale * 1e-4, scale * 2e-2, steps, dtype=np.float64)  # :26-34
        # create_gaussian_diffusion always builds a SpacedDiffusion (multimodal_script_util.py:223-242),
        # which re-derives betas from the base cumulative products even when every step is kept
        # (multimodal_respace.py:76-86): beta_i = 1 - abar_i / abar_{i-1}.  Differs from `base` in the
        # last ulp, so it is restated to keep the tables bit-identical.
        base_ac = np.cumprod(1.0 - base, axis=0)
        betas = np.array([1 - ac / prev for ac, prev in zip(base_ac, np.append(1.0, base_ac[:-1]))])
        self.betas = betas
        self.num_timesteps = steps
        alphas = 1.0 - betas
        ac = np.cumprod(alphas, axis=0)
        ac_prev = np.append(1.0, ac[:-1])
        self.alphas_cumprod = ac
        self.sqrt_alphas_cumprod = np.sqrt(ac)
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - ac)
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / ac)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / ac - 1.0)
        self.posterior_variance = betas * (1.0 - ac_prev) / (1.0 - ac)
        self.posterior_mean_coef1 = betas * np.sqrt(ac_prev) / (1.0 - ac)
        self.posterior_mean_coef2 = (1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac)
        # FIXED_LARGE (:292-305): variance = [posterior_variance[1], betas[1:]]
        self.fixed_large_log_variance = np.log(np.append(self.posterior_variance[1], betas[1:]))

    @staticmethod
    def _gather(arr, t, x):
        r = torch.from_numpy(arr)[t].float()
        return r.reshape(-1, *([1] * (x.dim() - 1)))

    def q_sample(self, x0, t, noise):  # :187-205
        return self._gather(self.sqrt_alphas_cumprod, t, x0) * x0 + \
            self._gather(self.sqrt_one_minus_alphas_cumprod, t, x0) * noise

    def p_sample_tail(self, x, eps, t, z, clip_denoised=True):
        """(:320-325, 345-350, 215-218, 453-470) -> (sample, pred_xstart)."""
        x0 = self._gather(self.sqrt_recip_alphas_cumprod, t, x) * x - self._gather(self.sqrt_recipm1_alphas_cumprod, t, x) * eps
        if clip_denoised:
            x0 = x0.clamp(-1, 1)
        mean = self._gather(self.posterior_mean_coef1, t, x) * x0 + self._gather(self.posterior_mean_coef2, t, x) * x
        logvar = self._gather(self.fixed_large_log_variance, t, x)
        nz = (t != 0).float().reshape(-1, *([1] * (x.dim() - 1)))
        return mean + nz * torch.exp(0.5 * logvar) * z, x0

    def p_sample(self, sd, cfg, x, t, noise, shifts, clip_denoised=True):
        """One p_sample step (:415-474) with injected noise dict; returns the reference's dict layout."""
        ev, ea = unet_forward(sd, cfg, x["video"], x["audio"], t, shifts)
        sv, x0v = self.p_sample_tail(x["video"], ev, t, noise["video"], clip_denoised)
        sa, x0a = self.p_sample_tail(x["audio"], ea, t, noise["audio"], clip_denoised)
        return {"sample": {"video": sv, "audio": sa}, "pred_start": {"video": x0v, "audio": x0a},
                "pred_noise": {"video": ev, "audio": ea}}

    def training_losses(self, sd, cfg, x_start, t, noise, shifts, dropout=None):
        """multimodal_training_losses (:1114-1203), EPSILON target, MSE: per-sample losses."""
        vt = self.q_sample(x_start["video"], t, noise["video"])
        at = self.q_sample(x_start["audio"], t, noise["audio"])
        ev, ea = unet_forward(sd, cfg, vt, at, t, shifts, dropout)
        mv = ((noise["video"] - ev) ** 2).flatten(1).mean(1)
        ma = ((noise["audio"] - ea) ** 2).flatten(1).mean(1)
        return {"loss": mv + ma, "mse_video": mv, "mse_audio": ma}


# --------------------------------------------------------------------------- synthetic weights
def param_shapes(cfg: UNetConfig) -> List[Tuple[str, Tuple[int, ...]]]:
    """(name, shape) of every parameter in the reference's registration order."""
    out: List[Tuple[str, Tuple[int, ...]]] = []
    mc = cfg.model_channels
    E = mc

    def lin(p, o, i):
        out.append((p + ".weight", (o, i)))
        out.append((p + ".bias", (o,)))

    def gn(p, c):
        out.append((p + ".GroupNorm.weight", (c,)))
        out.append((p + ".GroupNorm.bias", (c,)))

    def conv(p, o, i, *k):
        out.append((p + ".weight", (o, i, *k)))
        out.append((p + ".bias", (o,)))

    def vconv2d1d(p, i, o):
        conv(p + ".video_conv_spatial", o, i, 3, 3)
        conv(p + ".video_conv_temporal", o, o, 3)

    def attn(p, c):
        gn(p + ".norm", c)
        conv(p + ".qkv", 3 * c, c, 1)
        conv(p + ".proj_out", c, c, 1)

    def res(s: ResSpec):
        p = s.prefix
        gn(p + ".video_in_layers.0", s.cin)
        vconv2d1d(p + ".video_in_layers.2", s.cin, s.cout)
        gn(p + ".audio_in_layers.0", s.cin)
        conv(p + ".audio_in_layers.2.audio_conv", s.cout, s.cin, 3)
        lin(p + ".emb_layers.1", 2 * s.cout, E)
        gn(p + ".video_out_layers.0", s.cout)
        conv(p + ".video_out_layers.3.video_conv", s.cout, s.cout, 1, 1, 1)
        gn(p + ".audio_out_layers.0", s.cout)
        conv(p + ".audio_out_layers.3.audio_conv", s.cout, s.cout, 1)
        if s.cin != s.cout:
            conv(p + ".video_skip_connection.video_conv", s.cout, s.cin, 1, 1, 1)
            conv(p + ".audio_skip_connection.audio_conv", s.cout, s.cin, 1)
        if s.video_attention:
            attn(p + ".spatial_attention_block", s.cout)
            attn(p + ".temporal_attention_block", s.cout)
        if s.audio_attention:
            attn(p + ".audio_attention_block", s.cout)

    def cross(s: CrossSpec):
        p, c = s.prefix, s.channels
        gn(p + ".v_norm", c)
        gn(p + ".a_norm", c)
        conv(p + ".v_qkv", 3 * c, c, 1)
        conv(p + ".a_qkv", 3 * c, c, 1)
        conv(p + ".video_proj_out.video_conv", c, c, 1, 1, 1)
        conv(p + ".audio_proj_out.audio_conv", c, c, 1)

    lin("time_embed.0", E, mc)
    lin("time_embed.2", E, E)
    topo = build_topology(cfg)
    ch0 = int(cfg.channel_mult[0] * mc)
    for blk in topo.input_blocks + [topo.middle] + topo.output_blocks:
        for s in blk:
            if s == "initial":
                vconv2d1d("input_blocks.0.0.video_conv", cfg.video_size[1], ch0)
                conv("input_blocks.0.0.audio_conv.audio_conv", ch0, cfg.audio_size[0], 3)
            elif isinstance(s, ResSpec):
                res(s)
            else:
                cross(s)
    gn("audio_out.0", topo.final_ch)
    conv("audio_out.2.audio_conv", cfg.audio_out_channels, ch0, 3)
    gn("video_out.0", topo.final_ch)
    conv("video_out.2.video_conv", cfg.video_out_channels, ch0, 3, 3, 3)
    return out


def synthetic_state_dict(cfg: UNetConfig, seed: int = 0, std: float = 0.03) -> Dict[str, torch.Tensor]:
    """Deterministic random weights (no reference needed): N(0, std^2 * fan-in-ish) convs/linears,
    GroupNorm gain ~ 1, small biases.  Every tensor is non-zero so no branch is dead (the reference's
    zero_module init would make the output identically 0 — SURVEY.md App. D-1)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in param_shapes(cfg):
        if ".GroupNorm.weight" in name:
            sd[name] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith(".bias"):
            sd[name] = 0.05 * torch.randn(shape, generator=g)
        else:
            fan_in = int(np.prod(shape[1:]))
            sd[name] = torch.randn(shape, generator=g) * (1.0 / math.sqrt(fan_in)) * (std / 0.03)
    return sd
