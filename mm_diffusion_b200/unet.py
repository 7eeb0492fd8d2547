"""MultimodalUNet: drop-in for mm_diffusion.multimodal_unet.MultimodalUNet (reference
multimodal_unet.py:697-1101) whose forward is the hand-written sm_100a path in libmmdiff.so.

The module owns nn.Parameters with the reference's names, shapes and registration order
(SURVEY.md App. F), so checkpoints (`load_state_dict`, `load_state_dict_`), `.parameters()`,
EMA copies etc. behave like the reference's.  The network topology itself lives in the C
library (csrc/model.cu); this file only mirrors the Python surface.  There is no PyTorch
fallback: forward() requires a CUDA device and the built library.
"""
from __future__ import annotations

import ctypes as C
import math
import random
from typing import Dict, List, Sequence

import torch
import torch.nn as nn

from . import _lib
from ._lib import MmdConfig, MmdError, check

# parameters the reference zero-initialises (nn.py:141-147 zero_module call sites,
# multimodal_unet.py:275,377,385,609-610,1006,1011)
_ZERO_INIT_MARKERS = (
    ".video_out_layers.3.", ".audio_out_layers.3.", ".proj_out.", ".video_proj_out.", ".audio_proj_out.",
    "video_out.2.", "audio_out.2.",
)


def _fill_config(cfg: MmdConfig, video_size, audio_size, model_channels, video_out_channels, audio_out_channels,
                 num_res_blocks, cross_attention_resolutions, cross_attention_windows, cross_attention_shift,
                 video_attention_resolutions, audio_attention_resolutions, channel_mult, num_heads,
                 num_head_channels, max_batch):
    cfg.video_f, cfg.video_c, cfg.video_h, cfg.video_w = [int(x) for x in video_size]
    cfg.audio_c, cfg.audio_l = [int(x) for x in audio_size]
    cfg.model_channels = int(model_channels)
    cfg.video_out_channels = int(video_out_channels)
    cfg.audio_out_channels = int(audio_out_channels)
    cfg.num_res_blocks = int(num_res_blocks)

    def put(dst, values, what):
        values = [int(v) for v in values]
        if len(values) > _lib.MMD_MAX_LEVELS:
            raise ValueError(f"{what}: at most {_lib.MMD_MAX_LEVELS} entries")
        for i, v in enumerate(values):
            dst[i] = v
        return len(values)

    for m in channel_mult:
        if int(m) != m:
            raise ValueError("non-integer channel_mult is not supported by the sm_100a path")
    cfg.n_levels = put(cfg.channel_mult, channel_mult, "channel_mult")
    cfg.num_heads = int(num_heads)
    cfg.num_head_channels = int(num_head_channels)
    cfg.n_cross = put(cfg.cross_attention_resolutions, cross_attention_resolutions, "cross_attention_resolutions")
    put(cfg.cross_attention_windows, cross_attention_windows, "cross_attention_windows")
    if len(cross_attention_windows) != len(cross_attention_resolutions):
        raise ValueError("cross_attention_windows and cross_attention_resolutions differ in length")
    cfg.cross_attention_shift = int(bool(cross_attention_shift))
    cfg.n_video_attn = put(cfg.video_attention_resolutions, video_attention_resolutions, "video_attention_resolutions")
    cfg.n_audio_attn = put(cfg.audio_attention_resolutions, audio_attention_resolutions, "audio_attention_resolutions")
    cfg.max_batch = int(max_batch)


class _Node(nn.Module):
    """Anonymous container used to rebuild the reference's dotted parameter names."""


class _UNetFunction(torch.autograd.Function):
    """MultimodalUNet.forward under autograd: forward = mmd_model_forward_train (every intermediate kept on the
    device), backward = mmd_model_backward (hand-written sm_100a dgrad / wgrad / attention / GroupNorm adjoints).
    Parameters are passed as inputs so their gradients flow through autograd into the same nn.Parameter objects
    (DDP reducer hooks, MixedPrecisionTrainer and EMA of the reference's TrainLoop keep working)."""

    @staticmethod
    def forward(ctx, model, shifts, video, audio, timesteps, *params):
        vo, ao = model._run(video, audio, timesteps, shifts, train=True)
        ctx.model = model
        ctx.batch = video.shape[0]
        # the library keeps ONE tape per batch size: remember which forward this backward belongs to
        ctx.generation = int(_lib.load().mmd_model_train_generation(model._handle, ctx.batch))
        ctx.in_shapes = (tuple(video.shape), tuple(audio.shape))
        ctx.need_inputs = (video.requires_grad, audio.requires_grad)
        ctx.param_meta = [(p.requires_grad, tuple(p.shape)) for p in params]
        return vo, ao

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, d_vo, d_ao):
        model = ctx.model
        lib = _lib.load()
        dev = model._handle_device
        B = ctx.batch
        if int(lib.mmd_model_train_generation(model._handle, B)) != ctx.generation:
            raise MmdError("backward of a stale forward: another differentiable forward at this batch size ran after the "
                           "one being differentiated and overwrote its kept activations (run forward -> backward "
                           "pairs one at a time, or use torch.no_grad() for evaluation forwards)")
        if model.checkpoint_rng_compat:
            # The reference re-runs every (always checkpointed) CrossAttentionBlock in backward and draws a fresh
            # random.randint there (multimodal_unet.py:619-622 under nn.py:233-279), last block first (autograd visits
            # the checkpoints in reverse).  We differentiate the function that was evaluated, but consume the same
            # draws with the same bounds in the same order, so the global `random` stream stays aligned.
            model.draw_shifts(reverse=True)
        with torch.cuda.device(dev):
            vshape = (B, model._cfg.video_f, model.video_out_channels, model._cfg.video_h, model._cfg.video_w)
            ashape = (B, model.audio_out_channels, model._cfg.audio_l)
            d_vo = torch.zeros(vshape, dtype=torch.float32, device=dev) if d_vo is None else d_vo.to(torch.float32).contiguous()
            d_ao = torch.zeros(ashape, dtype=torch.float32, device=dev) if d_ao is None else d_ao.to(torch.float32).contiguous()
            n_flat = int(lib.mmd_model_param_floats(model._handle))
            d_vi = torch.empty(ctx.in_shapes[0], dtype=torch.float32, device=dev) if ctx.need_inputs[0] else None
            d_ai = torch.empty(ctx.in_shapes[1], dtype=torch.float32, device=dev) if ctx.need_inputs[1] else None
            wants_params = any(needs for needs, _ in ctx.param_meta)
            if model._flat_mode and wants_params:
                # flat-gradient mode: ONE persistent fp32 buffer owns every parameter gradient; .grad are views of it
                # (set here, not by autograd's AccumulateGrad, which would clone them).  A later backward without
                # zero_grad accumulates into the same buffer (micro-batches), like autograd does.
                buf, views = model._grad_buffer(dev, n_flat)
                fresh = model._plist[0].grad is None
                dst = buf if fresh else torch.empty(n_flat, dtype=torch.float32, device=dev)
                check(lib.mmd_model_backward(model._handle, B, d_vo.data_ptr(), d_ao.data_ptr(), dst.data_ptr(),
                                             _lib.ptr(d_vi), _lib.ptr(d_ai), _lib.current_stream_ptr()))
                if fresh:
                    for p, v, (needs, _) in zip(model._plist, views, ctx.param_meta):
                        p.grad = v if needs else None
                else:
                    buf.add_(dst)
                return (None, None, d_vi, d_ai, None, *([None] * len(ctx.param_meta)))
            flat = torch.empty(n_flat, dtype=torch.float32, device=dev) if wants_params else None
            check(lib.mmd_model_backward(model._handle, B, d_vo.data_ptr(), d_ao.data_ptr(), _lib.ptr(flat),
                                         _lib.ptr(d_vi), _lib.ptr(d_ai), _lib.current_stream_ptr()))
        grads = []
        for (needs, shape), off in zip(ctx.param_meta, model._param_offsets()):
            n = 1
            for d in shape:
                n *= d
            grads.append(flat[off:off + n].view(shape) if needs else None)
        # autograd mode: the gradients reach the nn.Parameters through AccumulateGrad (DDP reducer hooks and the
        # reference's MixedPrecisionTrainer see them as usual); `flat_grad` is this backward's buffer only
        model.flat_grad, model.flat_grad_views = flat, grads
        return (None, None, d_vi, d_ai, None, *grads)


class MultimodalUNet(nn.Module):
    """Same constructor signature as the reference (multimodal_unet.py:737-764)."""

    def __init__(self, video_size, audio_size, model_channels, video_out_channels, audio_out_channels, num_res_blocks,
                 cross_attention_resolutions, cross_attention_windows, cross_attention_shift,
                 video_attention_resolutions, audio_attention_resolutions, video_type="2d+1d", audio_type="1d",
                 dropout=0, channel_mult=(1, 2, 3, 4), num_classes=None, use_checkpoint=False, use_fp16=False,
                 num_heads=1, num_head_channels=-1, num_heads_upsample=-1, use_scale_shift_norm=False,
                 resblock_updown=True, max_batch=64):
        super().__init__()
        if video_type != "2d+1d" or audio_type != "1d":
            raise NotImplementedError("only video_type='2d+1d' / audio_type='1d' (the shipped configuration)")
        if num_classes is not None:
            raise NotImplementedError("class-conditional models are not supported (nor by the reference's scripts)")
        if not use_scale_shift_norm or not resblock_updown:
            # the reference's own non-default branches are broken
            raise NotImplementedError("use_scale_shift_norm=True and resblock_updown=True are required")
        self.video_size = video_size
        self.audio_size = audio_size
        self.model_channels = model_channels
        self.video_out_channels = video_out_channels
        self.audio_out_channels = audio_out_channels
        self.num_res_blocks = num_res_blocks
        self.cross_attention_resolutions = cross_attention_resolutions
        self.cross_attention_windows = cross_attention_windows
        self.cross_attention_shift = cross_attention_shift
        self.video_attention_resolutions = video_attention_resolutions
        self.audio_attention_resolutions = audio_attention_resolutions
        self.dropout = dropout
        self.channel_mult = channel_mult
        self.num_classes = num_classes
        self.use_checkpoint = use_checkpoint
        self.dtype = torch.float16 if use_fp16 else torch.float32
        self.num_heads = num_heads
        self.num_head_channels = num_head_channels
        self.num_heads_upsample = num_heads if num_heads_upsample == -1 else num_heads_upsample

        self._cfg = MmdConfig()
        _fill_config(self._cfg, video_size, audio_size, model_channels, video_out_channels, audio_out_channels,
                     num_res_blocks, cross_attention_resolutions, cross_attention_windows, cross_attention_shift,
                     video_attention_resolutions, audio_attention_resolutions, channel_mult, num_heads,
                     num_head_channels, max_batch)
        self.checkpoint_rng_compat = True   # see _UNetFunction.backward
        self._handle = None          # MmdModel* (created lazily on the parameters' CUDA device)
        self._handle_device = None
        self._synced: Dict[str, tuple] = {}
        self._synced_vsum = -1
        self._plist = None   # cached parameter list (the per-forward staleness check walks it)
        self._needs_sync = True
        self._drop_calls = 0
        self._flat_mode = False
        self._flat_views_ok = False
        self.flat_parameters = True    # keep the parameters as views of one flat fp32 buffer (see _flatten_parameters)
        self._flat_params = None
        self.flat_grad = None          # fp32 buffer holding every parameter gradient (see use_flat_gradients)
        self.flat_grad_views: List = []
        self._param_names: List[str] = []
        self._shift_bounds: List[int] = []
        self._build_parameters()

    # ------------------------------------------------------------------ construction
    def _inventory(self):
        """(name, shape) list + shift bounds from the C library's topology walk (host-only: works without a GPU)."""
        lib = _lib.load()
        h = C.c_void_p()
        check(lib.mmd_model_create(C.byref(self._cfg), C.byref(h)))
        try:
            n = lib.mmd_model_num_params(h)
            out = []
            name = C.c_char_p()
            ndim = C.c_int()
            shape = (C.c_int64 * 5)()
            for i in range(n):
                check(lib.mmd_model_param_info(h, i, C.byref(name), C.byref(ndim), shape))
                out.append((name.value.decode(), tuple(int(shape[j]) for j in range(ndim.value))))
            bounds = [lib.mmd_model_shift_bound(h, i) for i in range(lib.mmd_model_num_shifts(h))]
        finally:
            lib.mmd_model_destroy(h)
        return out, bounds

    def _build_parameters(self):
        names_shapes, bounds = self._inventory()
        self._shift_bounds = bounds
        for name, shape in names_shapes:
            parts = name.split(".")
            node = self
            for part in parts[:-1]:
                child = node._modules.get(part)
                if child is None:
                    child = _Node()
                    node.add_module(part, child)
                node = child
            p = nn.Parameter(torch.empty(shape, dtype=torch.float32))
            node.register_parameter(parts[-1], p)
            self._param_names.append(name)
        self.reset_parameters()

    @torch.no_grad()
    def reset_parameters(self):
        """PyTorch default Conv/Linear/GroupNorm initialisation + the reference's zero_module sites."""
        params = dict(self.named_parameters())
        for name, p in params.items():
            if any(mk in name for mk in _ZERO_INIT_MARKERS):
                p.zero_()
            elif ".GroupNorm." in name:
                p.fill_(1.0 if name.endswith("weight") else 0.0)
            elif name.endswith(".weight"):
                nn.init.kaiming_uniform_(p, a=math.sqrt(5))
            else:  # bias of a conv / linear: U(-1/sqrt(fan_in), 1/sqrt(fan_in))
                w = params[name[:-len("bias")] + "weight"]
                fan_in = w[0].numel()
                bound = 1.0 / math.sqrt(fan_in) if fan_in > 0 else 0.0
                p.uniform_(-bound, bound)
        self._needs_sync = True

    # ------------------------------------------------------------------ reference surface
    def convert_to_fp16(self):
        """Reference :1013-1021 casts conv weights to fp16 storage.  The sm_100a path always computes in fp16
        with fp32 accumulation from its own repacked copy, so only the output dtype changes here."""
        self.dtype = torch.float16

    def convert_to_fp32(self):
        self.dtype = torch.float32

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        out = super().load_state_dict(state_dict, strict=strict, assign=assign)
        self._needs_sync = True
        self._plist = None
        self._flat_views_ok = False
        return out

    def load_state_dict_(self, state_dict, is_strict=False):
        """Tolerant loader of the reference (:1033-1054): drop shape-mismatched keys, then load."""
        own = self.state_dict()
        for key, val in own.items():
            if key in state_dict and state_dict[key].shape != val.shape:
                state_dict.pop(key)
        self.load_state_dict(state_dict, strict=is_strict)

    def _apply(self, fn, recurse=True):
        out = super()._apply(fn, recurse)
        self._needs_sync = True
        self._plist = None
        return out

    def train(self, mode: bool = True):
        self._needs_sync = True
        return super().train(mode)

    # ------------------------------------------------------------------ device handle
    def _ensure_handle(self, device: torch.device):
        if self._handle is not None and self._handle_device == device:
            return
        lib = _lib.load()
        if self._handle is not None:
            lib.mmd_model_destroy(self._handle)
            self._handle = None
        import contextlib
        with (torch.cuda.device(device) if device.type == "cuda" else contextlib.nullcontext()):
            h = C.c_void_p()
            check(lib.mmd_model_create(C.byref(self._cfg), C.byref(h)))   # host-only until the first forward
        self._handle, self._handle_device = h, device
        self._synced = {}
        self._offsets = None
        self._needs_sync = True

    def _flat_params_ok(self) -> bool:
        """Every parameter is an fp32 view of self._flat_params at the library's offset (the layout of the flat gradient
        buffer)."""
        flat = self._flat_params
        if flat is None or self._plist is None:
            return False
        base = flat.data_ptr()
        for p, off in zip(self._plist, self._param_offsets()):
            if p.dtype != torch.float32 or p.data_ptr() != base + 4 * off:
                return False
        return True

    @torch.no_grad()
    def _flatten_parameters(self, device) -> bool:
        """Re-home the nn.Parameters (same objects, same names) as views of ONE flat fp32 device buffer laid out like the
        library's parameter arena.  An optimizer step then reaches the library as one device copy
        (mmd_model_set_params_flat) instead of one call per tensor, and the flat gradient buffer lines up element for
        element with it (fused multi-tensor updates are plain vector operations on the two buffers)."""
        if self._plist is None:
            self._plist = list(self.parameters())
        if any(p.dtype != torch.float32 or p.device != device for p in self._plist):
            return False   # e.g. module.half(): keep the per-tensor path
        n = int(_lib.load().mmd_model_param_floats(self._handle))
        flat = torch.zeros(n, dtype=torch.float32, device=device)
        for p, off in zip(self._plist, self._param_offsets()):
            view = flat[off:off + p.numel()].view(p.shape)
            view.copy_(p.detach())
            p.data = view
        self._flat_params = flat
        return True

    def flatten_for_training(self):
        """(flat fp32 parameter buffer, names, float offsets, shapes): the parameters re-homed as views of one buffer
        (see _flatten_parameters).  Used by fp16_util.MixedPrecisionTrainer, whose single master parameter IS this
        buffer, so the optimizer updates the model in place."""
        device = next(self.parameters()).device
        self._ensure_handle(device)
        self._plist = list(self.parameters())
        if not self._flat_params_ok() and not self._flatten_parameters(device):
            raise MmdError("flatten_for_training needs fp32 parameters on one device")
        return self._flat_params, list(self._param_names), list(self._param_offsets()), [tuple(p.shape) for p in self._plist]

    def mark_parameters_updated(self):
        """The flat parameter buffer was written directly (optimizer step on the master view): re-send and re-pack at the
        next forward."""
        self._needs_sync = True

    def _sync_parameters(self):
        lib = _lib.load()
        stream = _lib.current_stream_ptr()
        if self._plist is None:
            self._plist = list(self.parameters())
        if self.flat_parameters and (self._flat_params_ok() or self._flatten_parameters(self._handle_device)):
            check(lib.mmd_model_set_params_flat(self._handle, self._flat_params.data_ptr(), self._flat_params.numel(), stream))
            self._synced = {}
            self._needs_sync = False
            return
        for name, p in zip(self._param_names, self.parameters()):
            key = (p.data_ptr(), p._version, p.dtype)
            if self._synced.get(name) == key:
                continue
            src = p.detach()
            if src.dtype != torch.float32 or not src.is_contiguous():
                src = src.float().contiguous()
            check(lib.mmd_model_set_param(self._handle, name.encode(), src.data_ptr(), src.numel(), stream))
            self._synced[name] = key
        self._needs_sync = False

    def __getstate__(self):
        """copy.deepcopy / pickle: the library handle and everything derived from it stay behind; the copy creates its
        own on its first forward (EMA copies of a model, checkpoints of whole modules)."""
        state = dict(self.__dict__)
        state.update(_handle=None, _handle_device=None, _synced={}, _synced_vsum=-1, _needs_sync=True, _plist=None,
                     _offsets=None, _flat_params=None, flat_grad=None, flat_grad_views=[], _flat_views_ok=False)
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)

    def __del__(self):
        try:
            if self._handle is not None:
                _lib.load().mmd_model_destroy(self._handle)
        except Exception:
            pass

    @property
    def shift_bounds(self) -> List[int]:
        return list(self._shift_bounds)

    def draw_shifts(self, reverse: bool = False) -> List[int]:
        """One random.randint(0, F - window) per shifting cross-attention block, in execution order (reverse=True: last
        block first, the order of the reference's checkpoint recomputation in backward), from Python's global
        `random` — exactly the draws CrossAttentionBlock.attention_index makes (:619-622)."""
        bounds = list(reversed(self._shift_bounds)) if reverse else self._shift_bounds
        out = [random.randint(0, b) if b >= 0 else 0 for b in bounds]
        return list(reversed(out)) if reverse else out

    def use_flat_gradients(self, enable: bool = True):
        """Gradient hand-off of the training backward.
        False (default): parameter gradients flow through autograd into each nn.Parameter (DDP's reducer hooks, the
        reference's TrainLoop / MixedPrecisionTrainer work unchanged).
        True: the backward writes into ONE persistent flat fp32 buffer (`flat_grad`) and sets every `.grad` to a view of
        it — no per-parameter kernels, micro-batch accumulation in one add, and `parallel.allreduce_flat_gradients`
        exchanges all gradients in a single collective.  Not compatible with DistributedDataParallel (no autograd
        hooks fire for the parameters)."""
        self._flat_mode = bool(enable)
        return self

    def _grad_buffer(self, device, n_flat):
        if self._plist is None:
            self._plist = list(self.parameters())
        if self.flat_grad is None or self.flat_grad.numel() != n_flat or self.flat_grad.device != device or \
                not self._flat_views_ok:
            self.flat_grad = torch.zeros(n_flat, dtype=torch.float32, device=device)
            views = []
            for p, off in zip(self._plist, self._param_offsets()):
                views.append(self.flat_grad[off:off + p.numel()].view(p.shape))
            self.flat_grad_views = views
            self._flat_views_ok = True
        return self.flat_grad, self.flat_grad_views

    def _next_dropout_seed(self) -> int:
        """64-bit Philox seed of one training forward: a function of torch's CUDA seed (torch.manual_seed makes runs
        repeatable), the rank and a per-model call counter — no RNG state is consumed and nothing syncs."""
        self._drop_calls += 1
        rank = torch.distributed.get_rank() if (torch.distributed.is_available() and torch.distributed.is_initialized()) else 0
        x = (torch.cuda.initial_seed() + 0x9E3779B97F4A7C15 * (self._drop_calls + (rank << 40))) & 0xFFFFFFFFFFFFFFFF
        x ^= x >> 31
        return (x * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF

    def dropout_masks(self, batch: int):
        """Keep masks of the last training forward at this batch size, one (video [B,F,C,H,W], audio [B,C,L]) bool pair
        per ResBlock in execution order (test hook: the oracle applies the same masks)."""
        lib = _lib.load()
        n = lib.mmd_model_num_dropout_sites(self._handle, batch)
        out, pair = [], {}
        rows, ch, mod = C.c_int64(), C.c_int(), C.c_int()
        with torch.cuda.device(self._handle_device):
            for i in range(n):
                check(lib.mmd_model_dropout_site(self._handle, batch, i, C.byref(rows), C.byref(ch), C.byref(mod)))
                keep = torch.empty((rows.value, ch.value), dtype=torch.uint8, device=self._handle_device)
                check(lib.mmd_model_dropout_mask(self._handle, batch, i, keep.data_ptr(), _lib.current_stream_ptr()))
                if mod.value == 0:
                    hw = rows.value // (batch * self._cfg.video_f)
                    h = int(round((hw * self._cfg.video_h / self._cfg.video_w) ** 0.5))
                    pair["video"] = keep.view(batch, self._cfg.video_f, h, hw // h, ch.value).permute(0, 1, 4, 2, 3).bool()
                else:
                    pair["audio"] = keep.view(batch, rows.value // batch, ch.value).permute(0, 2, 1).bool()
                if len(pair) == 2:
                    out.append((pair["video"], pair["audio"]))
                    pair = {}
        return out

    def num_launches(self, batch: int) -> int:
        """Kernel launches of one forward at this batch size (0 until the plan has been built by a forward)."""
        return sum(s["kernels"] for s in self.plan_steps(batch))

    def plan_steps(self, batch: int):
        """[{kind, flops, bytes, kernels}] of the launch plan (algorithmic work per step)."""
        if self._handle is None:
            return []
        lib = _lib.load()
        n = lib.mmd_model_num_launches(self._handle, batch)
        out = []
        kind = C.c_char_p()
        fl, by, nk = C.c_double(), C.c_double(), C.c_int()
        for i in range(n):
            check(lib.mmd_model_step_info(self._handle, batch, i, C.byref(kind), C.byref(fl), C.byref(by), C.byref(nk)))
            out.append({"kind": kind.value.decode(), "flops": fl.value, "bytes": by.value, "kernels": nk.value})
        return out

    def profile(self, batch: int, reps: int = 3):
        """Per-step device time (ms, CUDA events on the current stream, un-graphed) of the plan for `batch`."""
        steps = self.plan_steps(batch)
        if not steps:
            raise MmdError("profile(): run a forward at this batch size first")
        buf = (C.c_float * len(steps))()
        with torch.cuda.device(self._handle_device):
            r = _lib.load().mmd_model_profile(self._handle, batch, reps, buf, len(steps), _lib.current_stream_ptr())
        if r < 0:
            check(r)
        for s, ms in zip(steps, buf):
            s["ms"] = float(ms)
        return steps

    # ------------------------------------------------------------------ forward
    def _param_offsets(self) -> List[int]:
        if getattr(self, "_offsets", None) is None:
            lib = _lib.load()
            self._offsets = [int(lib.mmd_model_param_offset(self._handle, i)) for i in range(len(self._param_names))]
        return self._offsets

    def num_backward_launches(self, batch: int) -> int:
        return 0 if self._handle is None else int(_lib.load().mmd_model_num_backward_launches(self._handle, batch))

    def profile_backward(self, batch: int, reps: int = 1):
        """[{kind, ms}] per backward step (un-graphed, CUDA events); needs a preceding differentiable forward."""
        lib = _lib.load()
        n = self.num_backward_launches(batch)
        buf = (C.c_float * max(n, 1))()
        with torch.cuda.device(self._handle_device):
            r = lib.mmd_model_profile_backward(self._handle, batch, reps, buf, n, _lib.current_stream_ptr())
        if r < 0:
            check(r)
        return [{"kind": lib.mmd_model_backward_step_kind(self._handle, batch, i).decode(), "ms": float(buf[i])} for i in range(n)]

    def _run(self, video, audio, timesteps, shifts, train: bool):
        """One library forward on detached fp32 inputs -> (video_out, audio_out) fp32."""
        device = next(self.parameters()).device
        B = video.shape[0]
        with torch.cuda.device(device):
            self._ensure_handle(device)
            # in-place updates that bypass load_state_dict / _apply / train() (an EMA swap through p.copy_(ema), an
            # optimizer step) bump the tensors' version counters: one cheap sum decides whether to look closer
            if self._plist is None:
                self._plist = list(self.parameters())
            vsum = sum(p._version for p in self._plist) + sum(p.data_ptr() for p in self._plist[:4])
            if self._needs_sync or self.training or train or vsum != self._synced_vsum:
                self._sync_parameters()
                self._synced_vsum = vsum
            if train:
                p_drop = float(self.dropout) if (self.training and self.dropout) else 0.0
                check(_lib.load().mmd_model_set_dropout(self._handle, p_drop, self._next_dropout_seed() if p_drop > 0 else 0))
            v = video.detach().to(torch.float32).contiguous()
            a = audio.detach().to(torch.float32).contiguous()
            t = timesteps.detach().to(device=device, dtype=torch.float32).contiguous()
            vo = torch.empty((B, self._cfg.video_f, self.video_out_channels, self._cfg.video_h, self._cfg.video_w),
                             dtype=torch.float32, device=device)
            ao = torch.empty((B, self.audio_out_channels, self._cfg.audio_l), dtype=torch.float32, device=device)
            n = len(self._shift_bounds)
            arr = (C.c_int32 * max(n, 1))(*[int(s) for s in shifts][:n])
            fn = _lib.load().mmd_model_forward_train if train else _lib.load().mmd_model_forward
            check(fn(self._handle, B, v.data_ptr(), a.data_ptr(), t.data_ptr(), arr, vo.data_ptr(), ao.data_ptr(),
                     _lib.current_stream_ptr()))
        return vo, ao

    def forward(self, video, audio, timesteps, label=None, shifts: Sequence[int] = None):
        """video [N,F,C,H,W], audio [N,C,L], timesteps [N] -> (video_out, audio_out) of self.dtype.

        `shifts` (optional) pins the random window shifts; by default they are drawn like the reference.
        Under autograd (parameters or inputs requiring grad, grad mode on) the call is differentiable: training
        (multimodal_gaussian_diffusion.py:1141) and gradient-guided conditional sampling (:815)."""
        assert (label is not None) == (self.num_classes is not None), \
            "must specify y if and only if the model is class-conditional"
        p0 = next(self.parameters())
        if not p0.is_cuda or not video.is_cuda:
            raise MmdError("MultimodalUNet.forward needs CUDA tensors: the denoising step is hand-written "
                           "sm_100a CUDA and has no CPU fallback")
        B = video.shape[0]
        if tuple(video.shape[1:]) != tuple(int(x) for x in self.video_size) or \
                tuple(audio.shape[1:]) != tuple(int(x) for x in self.audio_size) or audio.shape[0] != B:
            raise ValueError(f"input shapes {tuple(video.shape)} / {tuple(audio.shape)} do not match the model's "
                             f"video_size {self.video_size} / audio_size {self.audio_size}")
        if shifts is None:
            shifts = self.draw_shifts()
        params = list(self.parameters())
        differentiable = torch.is_grad_enabled() and (video.requires_grad or audio.requires_grad or
                                                      any(p.requires_grad for p in params))
        if differentiable:
            vo, ao = _UNetFunction.apply(self, list(shifts), video, audio, timesteps, *params)
        else:
            vo, ao = self._run(video, audio, timesteps, shifts, train=False)
        if self.dtype != torch.float32:
            vo, ao = vo.to(self.dtype), ao.to(self.dtype)
        return vo, ao
