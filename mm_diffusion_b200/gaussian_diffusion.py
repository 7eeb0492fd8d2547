"""Gaussian diffusion over {"video","audio"} dicts with a fused CUDA sampler tail.

API-compatible with the reference's mm_diffusion/multimodal_gaussian_diffusion.py (class / method /
attribute names and return layouts) but written for the B200 path: schedule tables are uploaded
to the device once (the reference re-uploads a numpy table ~10x per step, :1300), and the
per-step arithmetic after the model call (x0 prediction, clamp, posterior mean, noise injection)
is one kernel launch per modality through the C-ABI (mmd_p_sample_tail) instead of ~15 ATen ops.
"""
from __future__ import annotations

import enum
import math

import numpy as np
import torch as th

from . import _lib


def get_named_beta_schedule(schedule_name, num_diffusion_timesteps):
    """Reference :17-41."""
    if schedule_name == "linear":
        scale = 1000 / num_diffusion_timesteps
        return np.linspace(scale * 0.0001, scale * 0.02, num_diffusion_timesteps, dtype=np.float64)
    if schedule_name == "cosine":
        return betas_for_alpha_bar(num_diffusion_timesteps,
                                   lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2)
    raise NotImplementedError(f"unknown beta schedule: {schedule_name}")


def betas_for_alpha_bar(num_diffusion_timesteps, alpha_bar, max_beta=0.999):
    """Reference :44-61."""
    n = num_diffusion_timesteps
    return np.array([min(1 - alpha_bar((i + 1) / n) / alpha_bar(i / n), max_beta) for i in range(n)])


class ModelMeanType(enum.Enum):
    PREVIOUS_X = enum.auto()
    START_X = enum.auto()
    EPSILON = enum.auto()


class ModelVarType(enum.Enum):
    LEARNED = enum.auto()
    FIXED_SMALL = enum.auto()
    FIXED_LARGE = enum.auto()
    LEARNED_RANGE = enum.auto()


class LossType(enum.Enum):
    MSE = enum.auto()
    RESCALED_MSE = enum.auto()
    KL = enum.auto()
    RESCALED_KL = enum.auto()

    def is_vb(self):
        return self in (LossType.KL, LossType.RESCALED_KL)


def mean_flat(tensor):
    return tensor.mean(dim=list(range(1, tensor.dim())))


def _default_device():
    return th.device("cuda", th.cuda.current_device()) if th.cuda.is_available() else th.device("cpu")


class GaussianDiffusion:
    """Schedule tables + sampling / loss loops (reference :100-1286).  Supported model parameterisation:
    EPSILON or START_X mean, FIXED_LARGE / FIXED_SMALL variance (the shipped configuration is EPSILON +
    FIXED_LARGE, learn_sigma=False); learned-variance models raise NotImplementedError."""

    def __init__(self, *, betas, model_mean_type, model_var_type, loss_type, rescale_timesteps=False):
        self.model_mean_type = model_mean_type
        self.model_var_type = model_var_type
        self.loss_type = loss_type
        self.rescale_timesteps = rescale_timesteps
        if model_var_type in (ModelVarType.LEARNED, ModelVarType.LEARNED_RANGE):
            raise NotImplementedError("learn_sigma=True is outside the B200 hot path (production uses learn_sigma False)")
        if model_mean_type == ModelMeanType.PREVIOUS_X:
            raise NotImplementedError("PREVIOUS_X parameterisation is not supported")

        betas = np.array(betas, dtype=np.float64)
        assert betas.ndim == 1 and (betas > 0).all() and (betas <= 1).all()
        self.betas = betas
        self.num_timesteps = int(betas.shape[0])
        alphas = 1.0 - betas
        self.alphas_cumprod = np.cumprod(alphas, axis=0)
        self.alphas_cumprod_prev = np.append(1.0, self.alphas_cumprod[:-1])
        self.alphas_cumprod_next = np.append(self.alphas_cumprod[1:], 0.0)
        self.sqrt_alphas_cumprod = np.sqrt(self.alphas_cumprod)
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - self.alphas_cumprod)
        self.log_one_minus_alphas_cumprod = np.log(1.0 - self.alphas_cumprod)
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod - 1)
        self.posterior_variance = betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_log_variance_clipped = np.log(np.append(self.posterior_variance[1], self.posterior_variance[1:]))
        self.posterior_mean_coef1 = betas * np.sqrt(self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_mean_coef2 = (1.0 - self.alphas_cumprod_prev) * np.sqrt(alphas) / (1.0 - self.alphas_cumprod)
        if model_var_type == ModelVarType.FIXED_LARGE:  # reference :292-298
            self._model_variance = np.append(self.posterior_variance[1], betas[1:])
        else:
            self._model_variance = self.posterior_variance
        self._model_log_variance = np.log(np.append(self.posterior_variance[1], betas[1:])) \
            if model_var_type == ModelVarType.FIXED_LARGE else self.posterior_log_variance_clipped
        self._dev_tables = {}

    # ------------------------------------------------------------------ device tables
    def _tables(self, device):
        """[T, 8] fp32 on `device`: a=sqrt(1/abar), b=sqrt(1/abar-1), c1, c2, sigma, nonzero, sqrt(abar), sqrt(1-abar)."""
        key = str(device)
        tab = self._dev_tables.get(key)
        if tab is None:
            sigma = np.exp(0.5 * self._model_log_variance)
            nz = np.ones(self.num_timesteps)
            nz[0] = 0.0
            cols = [self.sqrt_recip_alphas_cumprod, self.sqrt_recipm1_alphas_cumprod, self.posterior_mean_coef1,
                    self.posterior_mean_coef2, sigma, nz, self.sqrt_alphas_cumprod, self.sqrt_one_minus_alphas_cumprod]
            tab = th.from_numpy(np.stack(cols, axis=1)).to(device=device, dtype=th.float32).contiguous()
            self._dev_tables[key] = tab
        return tab

    def _gather(self, arr, t, x):
        """fp32 table values at t, shaped for broadcasting against x (reference _extract_into_tensor :1289-1303)."""
        vals = th.from_numpy(np.asarray(arr, dtype=np.float64)).to(device=t.device)[t].float()
        return vals.reshape(-1, *([1] * (x.dim() - 1)))

    # ------------------------------------------------------------------ q(x_t | x_0)
    def q_mean_variance(self, x_start, t):
        return (self._gather(self.sqrt_alphas_cumprod, t, x_start) * x_start,
                self._gather(1.0 - self.alphas_cumprod, t, x_start).expand(x_start.shape),
                self._gather(self.log_one_minus_alphas_cumprod, t, x_start).expand(x_start.shape))

    def q_sample(self, x_start, t, noise=None):
        """sqrt(abar_t) x0 + sqrt(1-abar_t) eps (reference :187-205); one fused kernel on CUDA fp32 inputs."""
        if noise is None:
            noise = th.randn_like(x_start)
        assert noise.shape == x_start.shape
        if x_start.is_cuda and x_start.dtype == th.float32 and noise.dtype == th.float32:
            coef = self._tables(x_start.device)[t.long()][:, 6:8].contiguous()
            xs, nz = x_start.contiguous(), noise.contiguous()
            out = th.empty_like(xs)
            B = xs.shape[0]
            with th.cuda.device(xs.device):
                _lib.check(_lib.load().mmd_q_sample(xs.data_ptr(), nz.data_ptr(), coef.data_ptr(), B, xs.numel() // B,
                                                    out.data_ptr(), _lib.current_stream_ptr()))
            return out
        return self._gather(self.sqrt_alphas_cumprod, t, x_start) * x_start + \
            self._gather(self.sqrt_one_minus_alphas_cumprod, t, x_start) * noise

    def q_posterior_mean_variance(self, x_start, x_t, t):
        mean = self._gather(self.posterior_mean_coef1, t, x_t) * x_start + self._gather(self.posterior_mean_coef2, t, x_t) * x_t
        var = self._gather(self.posterior_variance, t, x_t).expand(x_t.shape)
        logvar = self._gather(self.posterior_log_variance_clipped, t, x_t).expand(x_t.shape)
        return mean, var, logvar

    def _predict_xstart_from_eps(self, x_t, t, eps):
        return self._gather(self.sqrt_recip_alphas_cumprod, t, x_t) * x_t - \
            self._gather(self.sqrt_recipm1_alphas_cumprod, t, x_t) * eps

    def _predict_eps_from_xstart(self, x_t, t, pred_xstart):
        return (self._gather(self.sqrt_recip_alphas_cumprod, t, x_t) * x_t - pred_xstart) / \
            self._gather(self.sqrt_recipm1_alphas_cumprod, t, x_t)

    def _scale_timesteps(self, t):
        return t.float() * (1000.0 / self.num_timesteps) if self.rescale_timesteps else t

    # ------------------------------------------------------------------ p(x_{t-1} | x_t)
    def p_mean_variance(self, model, x, t, clip_denoised=True, denoised_fn=None, model_kwargs=None):
        """Generic (unfused) statement of reference :231-343; the sampling loops use the fused tail instead."""
        model_kwargs = model_kwargs or {}
        B = x["video"].shape[0]
        assert t.shape == (B,)
        vo, ao = model(x["video"], x["audio"], self._scale_timesteps(t), **model_kwargs)
        out = {"mean": {}, "variance": {}, "log_variance": {}, "pred_xstart": {}, "model_predict": {"video": vo, "audio": ao}}
        for key, mo in (("video", vo), ("audio", ao)):
            xt = x[key]
            mo = mo.float()
            x0 = mo if self.model_mean_type == ModelMeanType.START_X else self._predict_xstart_from_eps(xt, t, mo)
            if denoised_fn is not None:
                x0 = denoised_fn(x0)
            if clip_denoised:
                x0 = x0.clamp(-1, 1)
            mean, _, _ = self.q_posterior_mean_variance(x0, xt, t)
            out["mean"][key] = mean
            out["variance"][key] = self._gather(self._model_variance, t, xt).expand(xt.shape)
            out["log_variance"][key] = self._gather(self._model_log_variance, t, xt).expand(xt.shape)
            out["pred_xstart"][key] = x0
        return out

    def _fusable(self, x, denoised_fn, cond_fn):
        if th.is_grad_enabled() and (x["video"].requires_grad or x["audio"].requires_grad):
            return False   # gradient-guided sampling differentiates through the tail: generic torch statement
        return (denoised_fn is None and cond_fn is None and self.model_mean_type == ModelMeanType.EPSILON and
                x["video"].is_cuda and x["video"].dtype == th.float32 and x["audio"].dtype == th.float32)

    def p_sample(self, model, x, t, clip_denoised=True, denoised_fn=None, cond_fn=None, model_kwargs=None, noise=None):
        """One ancestral step (reference :415-474).  Returns {"sample","pred_start","pred_noise"} dicts.
        `noise` (optional {"video","audio"}) injects the Gaussian draw; by default it is drawn with
        th.randn_like in the reference's order (video, then audio)."""
        model_kwargs = model_kwargs or {}
        if not self._fusable(x, denoised_fn, cond_fn):
            return self._p_sample_generic(model, x, t, clip_denoised, denoised_fn, cond_fn, model_kwargs, noise)
        vo, ao = model(x["video"], x["audio"], self._scale_timesteps(t), **model_kwargs)
        if isinstance(noise, dict) and "video" in noise and noise["video"].shape == x["video"].shape:
            zv, za = noise["video"], noise["audio"]
        else:
            zv = th.randn_like(x["video"])
            za = th.randn_like(x["audio"])
        coef = self._tables(t.device)[t.long()][:, :6].contiguous()
        lib = _lib.load()
        res = {"sample": {}, "pred_start": {}, "pred_noise": {"video": vo, "audio": ao}}
        B = t.shape[0]
        with th.cuda.device(x["video"].device):
            stream = _lib.current_stream_ptr()
            for key, eps, z in (("video", vo, zv), ("audio", ao, za)):
                xt = x[key].contiguous()
                e = eps.float().contiguous()
                zz = z.float().contiguous()
                sample = th.empty_like(xt)
                x0 = th.empty_like(xt)
                _lib.check(lib.mmd_p_sample_tail(xt.data_ptr(), e.data_ptr(), zz.data_ptr(), coef.data_ptr(), B,
                                                 xt.numel() // B, int(bool(clip_denoised)), sample.data_ptr(),
                                                 x0.data_ptr(), stream))
                res["sample"][key] = sample
                res["pred_start"][key] = x0
        return res

    def _p_sample_generic(self, model, x, t, clip_denoised, denoised_fn, cond_fn, model_kwargs, noise):
        out = self.p_mean_variance(model, x, t, clip_denoised=clip_denoised, denoised_fn=denoised_fn,
                                   model_kwargs=model_kwargs)
        if isinstance(noise, dict) and "video" in noise and noise["video"].shape == x["video"].shape:
            zv, za = noise["video"], noise["audio"]
        else:
            zv, za = th.randn_like(x["video"]), th.randn_like(x["audio"])
        if cond_fn is not None:
            raise NotImplementedError("cond_fn guidance is not part of the multimodal scripts")
        res = {"sample": {}, "pred_start": out["pred_xstart"], "pred_noise": out["model_predict"]}
        for key, z in (("video", zv), ("audio", za)):
            nzm = (t != 0).float().reshape(-1, *([1] * (x[key].dim() - 1)))
            res["sample"][key] = out["mean"][key] + nzm * th.exp(0.5 * out["log_variance"][key]) * z
        return res

    def _initial_noise(self, shape, device):
        # x_T is drawn on the CPU and moved, like the reference (:547-551), so seeds reproduce its samples
        v = th.randn(*shape["video"], device="cpu").to(device)
        a = th.randn(*shape["audio"], device="cpu").to(device)
        return {"video": v, "audio": a}

    def p_sample_loop(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None,
                      model_kwargs=None, device=None, progress=True):
        final = None
        for sample in self.p_sample_loop_progressive(model, shape, noise=noise, clip_denoised=clip_denoised,
                                                     denoised_fn=denoised_fn, cond_fn=cond_fn,
                                                     model_kwargs=model_kwargs, device=device, progress=progress):
            final = sample
        return final

    def p_sample_loop_progressive(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None,
                                  model_kwargs=None, device=None, progress=False):
        """Reference :523-582: T ancestral steps from x_T ~ N(0, I); yields every intermediate sample dict."""
        if device is None:
            device = _default_device()
        x = noise if isinstance(noise, dict) else self._initial_noise(shape, device)
        indices = list(range(self.num_timesteps))[::-1]
        if progress:
            from tqdm.auto import tqdm
            indices = tqdm(indices)
        B = shape["video"][0]
        for i in indices:
            t = th.full((B,), i, device=device, dtype=th.long)
            with th.no_grad():
                out = self.p_sample(model, x, t, clip_denoised=clip_denoised, denoised_fn=denoised_fn, cond_fn=cond_fn,
                                    model_kwargs=model_kwargs)
                yield out["sample"]
                x = out["sample"]

    # ------------------------------------------------------------------ zero-shot conditional sampling
    def conditional_p_sample_loop(self, model, shape, use_fp16, noise=None, clip_denoised=True, denoised_fn=None,
                                  cond_fn=None, model_kwargs=None, device=None, progress=True, class_scale=0.0):
        """Reference :584-639.  class_scale == 0 -> replacement method (:642-720); class_scale > 0 -> gradient
        guidance (:722-819), which backpropagates through the model to the target modality's input."""
        final = None
        loop = self.conditional_p_sample_loop_progressive_unscale if class_scale == 0 else \
            self.conditional_p_sample_loop_progressive_scale
        for sample in loop(
                model, shape, use_fp16, noise=noise, clip_denoised=clip_denoised, denoised_fn=denoised_fn,
                cond_fn=cond_fn, model_kwargs=model_kwargs, device=device, progress=progress, class_scale=class_scale):
            final = sample
        return final

    def conditional_p_sample_loop_progressive_unscale(self, model, shape, use_fp16, noise=None, clip_denoised=True,
                                                      denoised_fn=None, cond_fn=None, model_kwargs=None, device=None,
                                                      progress=False, class_scale=0.0):
        if device is None:
            device = _default_device()
        if noise is None:
            noise = self._initial_noise(shape, device)
        x = dict(noise)
        model_kwargs = model_kwargs if model_kwargs is not None else {}
        cond = {k: model_kwargs.pop(k) for k in ("video", "audio") if k in model_kwargs}
        indices = list(range(self.num_timesteps))[::-1]
        if progress:
            from tqdm.auto import tqdm
            indices = tqdm(indices)
        B = shape["video"][0]
        for i in indices:
            t = th.full((B,), i, device=device, dtype=th.long)
            for key, c in cond.items():  # overwrite the conditioned modality with q(x_t | condition) using the FIXED noise
                x[key] = self.q_sample(c.to(device=device, dtype=th.float32), t, noise=noise[key])
            with th.no_grad():
                out = self.p_sample(model, x, t, clip_denoised=clip_denoised, denoised_fn=denoised_fn, cond_fn=cond_fn,
                                    model_kwargs=model_kwargs)
                yield out["sample"]
                x = out["sample"]

    def conditional_p_sample_loop_progressive_scale(self, model, shape, use_fp16, noise=None, clip_denoised=True,
                                                    denoised_fn=None, cond_fn=None, model_kwargs=None, device=None,
                                                    progress=False, class_scale=3.0):
        """Gradient-guided zero-shot conditional sampling (reference :722-819): each step overwrites the conditioned
        modality with q(x_t | condition) under the fixed noise, takes one ancestral step with the target modality's
        input requiring grad, and moves the target by -class_scale * sqrt(alpha_bar_i) * d/dx_target of the MSE between
        the predicted and the true x_{t-1} of the conditioned modality.  Quirks kept: the 2^20 loss scale of the fp16
        path is not divided out again (:813-816), and t - 1 wraps to the last table entry at t = 0 (masked anyway)."""
        if device is None:
            device = _default_device()
        if noise is None:
            noise = self._initial_noise(shape, device)
        x = dict(noise)
        model_kwargs = model_kwargs if model_kwargs is not None else {}
        cond = {k: model_kwargs.pop(k) for k in ("video", "audio") if k in model_kwargs}
        if len(cond) != 1:
            raise ValueError("gradient-guided sampling needs exactly one of model_kwargs['video'] / ['audio'] as the condition")
        condition = next(iter(cond))
        target = "audio" if condition == "video" else "video"
        c = cond[condition].to(device=device, dtype=th.float32)
        indices = list(range(self.num_timesteps))[::-1]
        if progress:
            from tqdm.auto import tqdm
            indices = tqdm(indices)
        B = shape["video"][0]
        sqrt_ab = np.sqrt(self.alphas_cumprod)
        for i in indices:
            t = th.full((B,), i, device=device, dtype=th.long)
            with th.no_grad():
                x[condition] = self.q_sample(c, t, noise=noise[condition])
                prev_cond = self.q_sample(c, (t - 1) % self.num_timesteps, noise=noise[condition])
            with th.enable_grad():
                nzm = (t != 0).float().reshape(-1, *([1] * (x[target].dim() - 1)))
                x[target] = x[target].detach().requires_grad_()
                out = self.p_sample(model, x, t, clip_denoised=clip_denoised, denoised_fn=denoised_fn, cond_fn=cond_fn,
                                    model_kwargs=model_kwargs)
                pred = out["sample"]
                loss = mean_flat((pred[condition] - prev_cond) ** 2)
                loss_scale = float(2 ** 20) if use_fp16 else 1.0
                grad = th.autograd.grad(loss.mean() * loss_scale, x[target])[0]
                # like the reference, the yielded dict keeps q(x_t | condition) for the conditioned modality (:760-770, 816)
                x = {condition: x[condition],
                     target: (pred[target] - nzm * grad * class_scale * float(sqrt_ab[i])).detach()}
            yield x

    # ------------------------------------------------------------------ DDIM
    def ddim_sample(self, model, x, t, clip_denoised=True, denoised_fn=None, cond_fn=None, model_kwargs=None, eta=0.0):
        """Reference :821-901 (deterministic for eta = 0)."""
        out = self.p_mean_variance(model, x, t, clip_denoised=clip_denoised, denoised_fn=denoised_fn,
                                   model_kwargs=model_kwargs)
        res = {"sample": {}, "pred_xstart": out["pred_xstart"]}
        for key in ("video", "audio"):
            xt = x[key]
            x0 = out["pred_xstart"][key]
            eps = self._predict_eps_from_xstart(xt, t, x0)
            ab = self._gather(self.alphas_cumprod, t, xt)
            ab_prev = self._gather(self.alphas_cumprod_prev, t, xt)
            sigma = eta * th.sqrt((1 - ab_prev) / (1 - ab)) * th.sqrt(1 - ab / ab_prev)
            mean = x0 * th.sqrt(ab_prev) + th.sqrt(1 - ab_prev - sigma ** 2) * eps
            nzm = (t != 0).float().reshape(-1, *([1] * (xt.dim() - 1)))
            res["sample"][key] = mean + nzm * sigma * th.randn_like(xt)
        return res

    def ddim_sample_loop(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None,
                         model_kwargs=None, device=None, progress=False, eta=0.0):
        if device is None:
            device = _default_device()
        x = noise if isinstance(noise, dict) else self._initial_noise(shape, device)
        indices = list(range(self.num_timesteps))[::-1]
        if progress:
            from tqdm.auto import tqdm
            indices = tqdm(indices)
        B = shape["video"][0]
        for i in indices:
            t = th.full((B,), i, device=device, dtype=th.long)
            with th.no_grad():
                x = self.ddim_sample(model, x, t, clip_denoised=clip_denoised, denoised_fn=denoised_fn, cond_fn=cond_fn,
                                     model_kwargs=model_kwargs, eta=eta)["sample"]
        return x

    # ------------------------------------------------------------------ training objective
    def multimodal_training_losses(self, model, x_start, t, model_kwargs=None, noise=None):
        """eps-MSE per modality (reference :1114-1203): {"loss","mse_video","mse_audio"} of shape [N]; differentiable
        wrt the model parameters through the sm_100a backward (unet._UNetFunction)."""
        model_kwargs = model_kwargs or {}
        if noise is None:
            noise = {"video": th.randn_like(x_start["video"]), "audio": th.randn_like(x_start["audio"])}
        vt = self.q_sample(x_start["video"], t, noise=noise["video"])
        at = self.q_sample(x_start["audio"], t, noise=noise["audio"])
        vo, ao = model(vt, at, self._scale_timesteps(t), **model_kwargs)
        target_v = noise["video"] if self.model_mean_type == ModelMeanType.EPSILON else x_start["video"]
        target_a = noise["audio"] if self.model_mean_type == ModelMeanType.EPSILON else x_start["audio"]
        terms = {"mse_video": mean_flat((target_v - vo.float()) ** 2), "mse_audio": mean_flat((target_a - ao.float()) ** 2)}
        terms["loss"] = terms["mse_video"] + terms["mse_audio"]
        return terms
