"""Timestep respacing (API of the reference's mm_diffusion/multimodal_respace.py)."""
from __future__ import annotations

import numpy as np
import torch as th

from .gaussian_diffusion import GaussianDiffusion


def space_timesteps(num_timesteps, section_counts):
    """Which of the original steps to keep: "ddimN" = fixed stride giving exactly N steps, otherwise a list /
    comma string of per-section counts spread evenly over equal sections (reference :6-59)."""
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            want = int(section_counts[4:])
            for stride in range(1, num_timesteps):
                if len(range(0, num_timesteps, stride)) == want:
                    return set(range(0, num_timesteps, stride))
            raise ValueError(f"cannot create exactly {num_timesteps} steps with an integer stride")
        section_counts = [int(x) for x in section_counts.split(",")]
    base, extra = divmod(num_timesteps, len(section_counts))
    start, kept = 0, []
    for i, count in enumerate(section_counts):
        size = base + (1 if i < extra else 0)
        if size < count:
            raise ValueError(f"cannot divide section of {size} steps into {count}")
        stride = 1 if count <= 1 else (size - 1) / (count - 1)
        pos = 0.0
        for _ in range(count):
            kept.append(start + round(pos))
            pos += stride
        start += size
    return set(kept)


class SpacedDiffusion(GaussianDiffusion):
    """Diffusion over a subset of the base steps (reference :62-124).  Betas are re-derived from the base
    cumulative products, beta_i = 1 - abar_i / abar_prev(kept) — also when every step is kept, which is why the
    tables differ from the plain linear schedule in the last ulp."""

    def __init__(self, use_timesteps, **kwargs):
        self.use_timesteps = set(use_timesteps)
        self.timestep_map = []
        self.original_num_steps = len(kwargs["betas"])
        base_ac = np.cumprod(1.0 - np.array(kwargs["betas"], dtype=np.float64), axis=0)
        last, new_betas = 1.0, []
        for i, ac in enumerate(base_ac):
            if i in self.use_timesteps:
                new_betas.append(1 - ac / last)
                last = ac
                self.timestep_map.append(i)
        kwargs["betas"] = np.array(new_betas)
        super().__init__(**kwargs)
        self._map_cache = {}

    def _wrap_model(self, model):
        if isinstance(model, _WrappedModel):
            return model
        return _WrappedModel(model, self.timestep_map, self.rescale_timesteps, self.original_num_steps, self._map_cache)

    def p_mean_variance(self, model, *args, **kwargs):
        return super().p_mean_variance(self._wrap_model(model), *args, **kwargs)

    def p_sample(self, model, *args, **kwargs):
        return super().p_sample(self._wrap_model(model), *args, **kwargs)

    def multimodal_training_losses(self, model, *args, **kwargs):
        return super().multimodal_training_losses(self._wrap_model(model), *args, **kwargs)

    def _scale_timesteps(self, t):
        return t  # done by the wrapped model


class _WrappedModel:
    """Maps spaced indices back to original timesteps before calling the model (reference :127-139); the map
    tensor is cached per device instead of being rebuilt every call."""

    def __init__(self, model, timestep_map, rescale_timesteps, original_num_steps, cache=None):
        self.model = model
        self.timestep_map = timestep_map
        self.rescale_timesteps = rescale_timesteps
        self.original_num_steps = original_num_steps
        self._cache = cache if cache is not None else {}

    def __call__(self, video_x, audio_x, ts, **kwargs):
        key = (str(ts.device), ts.dtype)
        mt = self._cache.get(key)
        if mt is None:
            mt = th.tensor(self.timestep_map, device=ts.device, dtype=ts.dtype)
            self._cache[key] = mt
        new_ts = mt[ts]
        if self.rescale_timesteps:
            new_ts = new_ts.float() * (1000.0 / self.original_num_steps)
        return self.model(video_x, audio_x, new_ts, **kwargs)
