"""Optimizer tail of the training step on flat buffers: drop-in for the reference's `MixedPrecisionTrainer`
(mm_diffusion/fp16_util.py:142-245) and `update_ema` (mm_diffusion/nn.py:128-138) for models whose parameters and
gradients already live in ONE flat fp32 buffer each (mm_diffusion_b200.unet.MultimodalUNet).

What the reference does per optimizer step with `use_fp16=True` (its shipped flag), and what replaces it here:

  reference (fp16_util.py)                                     here
  -----------------------------------------------------------  -----------------------------------------------------
  model_grads_to_master_grads: flatten 1046 .grad tensors      nothing: the backward wrote the flat gradient buffer and
    into the two master gradients (:51-61)                       it IS the master gradient
  _compute_norms: one th.norm(...).item() host sync per         two fused reductions over the flat buffers, ONE host sync
    master tensor for parameters and gradients (:228-236;
    2 x 1046 syncs when use_fp16=False)
  p.grad.mul_(1 / loss_scale) per master tensor (:217-218)      one in-place scale of the flat gradient
  opt.step() on the master tensors                              opt.step() on ONE flat nn.Parameter (a handful of kernels)
  master_params_to_model_params: 1046 copies back (:64-74)      nothing: the model's parameters are views of the master
  zero_grad: 1046 .grad.zero_() (:127-133)                      the .grad views are dropped; the next backward rewrites
                                                                the flat buffer
  update_ema per tensor (nn.py:137-138)                         unchanged call, but on one flat tensor per EMA rate

The class keeps the reference's constructor, attributes (`model`, `use_fp16`, `master_params`, `model_params`,
`lg_loss_scale`, `fp16_scale_growth`) and methods (`zero_grad`, `backward`, `optimize`, `master_params_to_state_dict`,
`state_dict_to_master_params`), so a training loop written against the reference's trainer
(multimodal_train_util.py:267-278, 332-334, 463-482: zero_grad / backward / optimize / update_ema / checkpoint dicts)
drives it unchanged, and `compat.install(optimizer_tail=True)` aliases this module as `mm_diffusion.fp16_util`.
Checkpoints are interchangeable: `master_params_to_state_dict` returns the reference's `{name: tensor}` schema.

Multi-GPU: the gradients never pass through autograd's per-parameter hooks in this mode, so the model must NOT be wrapped
in DistributedDataParallel (the reference's TrainLoop wraps unconditionally, :127-136 — keep the default trainer there);
`optimize()` itself all-reduces the flat gradient buffer (one NCCL collective) when a process group is up.
"""
from __future__ import annotations

import math

import torch as th
import torch.nn as nn

INITIAL_LOG_LOSS_SCALE = 20.0


def convert_module_to_f16(l):   # noqa: E741  (reference name; fp16_util.py:13-20)
    """Reference surface: casts Conv weights to half.  The sm_100a model keeps fp32 nn.Parameters (its fp16 packs are
    the library's own), so this only touches plain torch conv modules."""
    if isinstance(l, (nn.Conv1d, nn.Conv2d, nn.Conv3d)):
        l.weight.data = l.weight.data.half()
        if l.bias is not None:
            l.bias.data = l.bias.data.half()


def convert_module_to_f32(l):   # noqa: E741  (fp16_util.py:23-30)
    if isinstance(l, (nn.Conv1d, nn.Conv2d, nn.Conv3d)):
        l.weight.data = l.weight.data.float()
        if l.bias is not None:
            l.bias.data = l.bias.data.float()


def check_overflow(value):
    return (value == float("inf")) or (value == -float("inf")) or (value != value)


def update_ema(target_params, source_params, rate=0.99):
    """nn.py:128-138, unchanged semantics: targ <- rate * targ + (1 - rate) * src, one fused lerp per tensor (the flat
    trainer hands it one tensor per EMA rate)."""
    for targ, src in zip(target_params, source_params):
        targ.detach().lerp_(src.detach(), 1.0 - rate)


def _logger():
    """The reference's logger when its package is importable (TrainLoop reads the values back from it), else a no-op."""
    try:
        from mm_diffusion import logger  # type: ignore
        return logger
    except Exception:  # noqa: BLE001
        class _Null:
            @staticmethod
            def logkv(*a, **k):
                pass
            logkv_mean = log = logkv
        return _Null


class MixedPrecisionTrainer:
    """Same contract as the reference's class; `model` must be a mm_diffusion_b200 MultimodalUNet on a CUDA device (or
    any module exposing `flatten_for_training()` -> (flat_params, names, offsets, shapes) and `flat_grad`)."""

    def __init__(self, *, model, use_fp16=False, fp16_scale_growth=1e-3, initial_lg_loss_scale=INITIAL_LOG_LOSS_SCALE):
        self.model = model
        self.use_fp16 = use_fp16
        self.fp16_scale_growth = fp16_scale_growth
        self.lg_loss_scale = initial_lg_loss_scale
        self.model_params = list(model.parameters())
        flat, self._names, self._offsets, self._shapes = model.flatten_for_training()
        # ONE master parameter: the flat buffer the model's own parameters are views of (no second copy to keep in sync)
        self._master = nn.Parameter(flat, requires_grad=True)
        self.master_params = [self._master]
        self.param_groups_and_shapes = None
        if use_fp16:
            model.convert_to_fp16()
        model.use_flat_gradients(True)

    # ------------------------------------------------------------------ reference methods
    def zero_grad(self):
        for p in self.model_params:
            p.grad = None          # the next backward rewrites the flat gradient buffer from scratch
        self._master.grad = None

    def backward(self, loss: th.Tensor):
        if self.use_fp16:
            (loss * (2 ** self.lg_loss_scale)).backward()
        else:
            loss.backward()

    def _norms(self, grad_scale=1.0):
        sq = th.stack([th.linalg.vector_norm(self.model.flat_grad, dtype=th.float32) ** 2,
                       th.linalg.vector_norm(self._master.detach(), dtype=th.float32) ** 2])
        gn2, pn2 = sq.tolist()   # the step's one host sync
        return (math.sqrt(gn2) / grad_scale if gn2 == gn2 else float("nan")), math.sqrt(pn2)

    def optimize(self, opt: th.optim.Optimizer):
        logger = _logger()
        if self.model.flat_grad is None or self.model_params[0].grad is None:
            raise RuntimeError("MixedPrecisionTrainer.optimize: no backward has run since zero_grad()")
        scale = 2 ** self.lg_loss_scale if self.use_fp16 else 1.0
        if self.use_fp16:
            logger.logkv_mean("lg_loss_scale", self.lg_loss_scale)
        if th.distributed.is_available() and th.distributed.is_initialized() and th.distributed.get_world_size() > 1:
            from .parallel import allreduce_flat_gradients
            allreduce_flat_gradients(self.model)   # the data-parallel exchange: one collective over the flat buffer
        grad_norm, param_norm = self._norms(grad_scale=scale)
        if self.use_fp16 and check_overflow(grad_norm):
            self.lg_loss_scale -= 1
            logger.log(f"Found NaN, decreased lg_loss_scale to {self.lg_loss_scale}")
            self.zero_grad()
            return False
        logger.logkv("current_grad_norm", grad_norm)
        logger.logkv("current_param_norm", param_norm)
        logger.logkv_mean("grad_norm", grad_norm)
        logger.logkv_mean("param_norm", param_norm)
        g = self.model.flat_grad
        if self.use_fp16:
            g.mul_(1.0 / scale)
        self._master.grad = g
        opt.step()
        self._master.grad = None
        self.model.mark_parameters_updated()
        if self.use_fp16:
            self.lg_loss_scale += self.fp16_scale_growth
        return True

    # ------------------------------------------------------------------ checkpoints (reference schema)
    def _unflatten(self, flat):
        flat = flat.detach().reshape(-1)
        return {n: flat[o:o + math.prod(s)].view(s) for n, o, s in zip(self._names, self._offsets, self._shapes)}

    def master_params_to_state_dict(self, master_params):
        state_dict = self.model.state_dict()
        for name, value in self._unflatten(master_params[0]).items():
            assert name in state_dict
            state_dict[name] = value
        return state_dict

    def state_dict_to_master_params(self, state_dict):
        flat = th.zeros_like(self._master.detach())
        for name, view in self._unflatten(flat).items():
            view.copy_(state_dict[name])
        return [nn.Parameter(flat, requires_grad=True)]
