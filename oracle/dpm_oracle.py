"""CPU oracle for the multimodal DPM-Solver driver.  TEST INFRASTRUCTURE ONLY (see oracle/mmdiff_oracle.py header:
only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this).

Restates ``mm_diffusion/multimodal_dpm_solver_plus.py`` (reference repo) the way the reference evaluates it — fp32
torch tensors, one time value per batch element broadcast over the sample, per-modality tensor arithmetic — so that the
product (host scalars + fused CUDA lincomb) is checked against an independently structured statement.  Pinned against
tests/golden/dpm_small.pt, which oracle/make_golden_dpm.py produced by running the unmodified reference.

The ``model`` argument is any callable ``model(video, audio, t_int[B]) -> (video_eps, audio_eps)``.
"""
from __future__ import annotations

import torch

KEYS = ("video", "audio")


def piecewise_linear(x, xp, yp):
    """f(x) through keypoints (xp ascending), extrapolating with the outermost segments (reference :1306-1346)."""
    k = xp.numel()
    x = x.contiguous()
    j = torch.bucketize(x, xp).clamp(1, k - 1)          # right end of the segment
    x0, x1, y0, y1 = xp[j - 1], xp[j], yp[j - 1], yp[j]
    return y0 + (x - x0) * (y1 - y0) / (x1 - x0)


class Schedule:
    """Discrete VP schedule: t_n = (n+1)/N, log alpha_t linear between trained steps (reference :77-180)."""

    def __init__(self, alphas_cumprod):
        self.log_alpha = 0.5 * torch.log(torch.as_tensor(alphas_cumprod, dtype=torch.float32))
        self.N = self.log_alpha.numel()
        self.t = torch.linspace(0.0, 1.0, self.N + 1)[1:]
        self.T = 1.0

    def log_mean(self, t):                               # :131-142
        return piecewise_linear(t.reshape(-1), self.t, self.log_alpha)

    def alpha(self, t):
        return torch.exp(self.log_mean(t))

    def sigma(self, t):                                  # :150-154
        return torch.sqrt(1.0 - torch.exp(2.0 * self.log_mean(t)))

    def lam(self, t):                                    # :156-162
        lm = self.log_mean(t)
        return lm - 0.5 * torch.log(1.0 - torch.exp(2.0 * lm))

    def inv_lam(self, lamb):                             # :164-180
        la = -0.5 * torch.logaddexp(torch.zeros(1), -2.0 * lamb.reshape(-1))
        return piecewise_linear(la, torch.flip(self.log_alpha, [0]), torch.flip(self.t, [0]))


def _bc(v, x):
    return v.reshape((-1,) + (1,) * (x.dim() - 1))


class DPMOracle:
    def __init__(self, model, alphas_cumprod, predict_x0=False, thresholding=False, max_val=1.0):
        self.model, self.ns = model, Schedule(alphas_cumprod)
        self.predict_x0, self.thresholding, self.max_val = predict_x0, thresholding, max_val
        self.model_times = []

    # ---- model functions (:285-312, :413-449)
    def eps(self, x, t):
        b = x["video"].shape[0]
        t = t.reshape(-1).expand(b) if t.numel() == 1 else t
        t_in = ((t - 1.0 / self.ns.N) * self.ns.N).to(torch.int)
        self.model_times.append(t_in.clone())
        v, a = self.model(x["video"], x["audio"], t_in)
        return {"video": v, "audio": a}

    def x0(self, x, t):
        e = self.eps(x, t)
        al, sg = self.ns.alpha(t), self.ns.sigma(t)
        out = {}
        for k in KEYS:
            v = (x[k] - _bc(sg, x[k]) * e[k]) / _bc(al, x[k])
            if self.thresholding:
                s = torch.quantile(v.abs().reshape(v.shape[0], -1), 0.995, dim=1)
                s = _bc(torch.maximum(s, torch.ones_like(s)), v)
                v = torch.clamp(v, -s, s) / (s / self.max_val)
            out[k] = v
        return out

    def f(self, x, t):
        return self.x0(x, t) if self.predict_x0 else self.eps(x, t)

    # ---- updates
    def first(self, x, s, t, m_s=None, ret=False):       # :532-588
        ns = self.ns
        h = ns.lam(t) - ns.lam(s)
        if m_s is None:
            m_s = self.f(x, s)
        out = {}
        for k in KEYS:
            if self.predict_x0 or k == "audio":          # audio always takes this form (:577-580)
                phi = torch.expm1(-h) if self.predict_x0 else torch.expm1(h)
                out[k] = _bc(ns.sigma(t) / ns.sigma(s), x[k]) * x[k] - _bc(ns.alpha(t) * phi, x[k]) * m_s[k]
            else:
                out[k] = (_bc(torch.exp(ns.log_mean(t) - ns.log_mean(s)), x[k]) * x[k]
                          - _bc(ns.sigma(t) * torch.expm1(h), x[k]) * m_s[k])
        return (out, {"m_s": m_s}) if ret else out

    def _lead(self, s, t):
        """(coefficient of x, coefficient scale of the model terms, signed h) for the two parametrisations."""
        ns = self.ns
        h = ns.lam(t) - ns.lam(s)
        if self.predict_x0:
            return ns.sigma(t) / ns.sigma(s), ns.alpha(t), -h
        return torch.exp(ns.log_mean(t) - ns.log_mean(s)), ns.sigma(t), h

    def second(self, x, s, t, r1=0.5, m_s=None, ret=False, solver_type="dpm_solver"):   # :590-704
        ns = self.ns
        r1 = 0.5 if r1 is None else r1
        h = ns.lam(t) - ns.lam(s)
        s1 = ns.inv_lam(ns.lam(s) + r1 * h)
        if m_s is None:
            m_s = self.f(x, s)
        a1, b1, g1 = self._lead(s, s1)
        x1 = {k: _bc(a1, x[k]) * x[k] - _bc(b1 * torch.expm1(g1), x[k]) * m_s[k] for k in KEYS}
        m_s1 = self.f(x1, s1)
        a, b, g = self._lead(s, t)
        out = {}
        for k in KEYS:
            base = _bc(a, x[k]) * x[k] - _bc(b * torch.expm1(g), x[k]) * m_s[k]
            diff = m_s1[k] - m_s[k]
            if solver_type == "dpm_solver":
                out[k] = base - (0.5 / r1) * _bc(b * torch.expm1(g), x[k]) * diff
            elif self.predict_x0:
                out[k] = base + (1.0 / r1) * _bc(b * ((torch.exp(-h) - 1.0) / h + 1.0), x[k]) * diff
            else:
                out[k] = base - (1.0 / r1) * _bc(b * ((torch.exp(h) - 1.0) / h - 1.0), x[k]) * diff
        return (out, {"m_s": m_s, "m_s1": m_s1}) if ret else out

    def third(self, x, s, t, r1=1.0 / 3.0, r2=2.0 / 3.0, m_s=None, m_s1=None, ret=False):   # :706-887 ('dpm_solver')
        ns = self.ns
        r1 = 1.0 / 3.0 if r1 is None else r1
        r2 = 2.0 / 3.0 if r2 is None else r2
        h = ns.lam(t) - ns.lam(s)
        s1, s2 = ns.inv_lam(ns.lam(s) + r1 * h), ns.inv_lam(ns.lam(s) + r2 * h)
        sgn = -1.0 if self.predict_x0 else 1.0           # phi functions are taken at -h for the data model
        phi_1 = torch.expm1(sgn * h)
        phi_22 = torch.expm1(sgn * r2 * h) / (r2 * h) - sgn
        phi_2 = phi_1 / h - sgn
        if m_s is None:
            m_s = self.f(x, s)
        if m_s1 is None:
            a1, b1, g1 = self._lead(s, s1)
            m_s1 = self.f({k: _bc(a1, x[k]) * x[k] - _bc(b1 * torch.expm1(g1), x[k]) * m_s[k] for k in KEYS}, s1)
        a2, b2, g2 = self._lead(s, s2)
        x2 = {k: _bc(a2, x[k]) * x[k] - _bc(b2 * torch.expm1(g2), x[k]) * m_s[k]
              - sgn * (r2 / r1) * _bc(b2 * phi_22, x[k]) * (m_s1[k] - m_s[k]) for k in KEYS}
        m_s2 = self.f(x2, s2)
        a, b, _ = self._lead(s, t)
        out = {k: _bc(a, x[k]) * x[k] - _bc(b * phi_1, x[k]) * m_s[k]
               - sgn * (1.0 / r2) * _bc(b * phi_2, x[k]) * (m_s2[k] - m_s[k]) for k in KEYS}
        return (out, {"m_s": m_s, "m_s1": m_s1, "m_s2": m_s2}) if ret else out

    def multi2(self, x, m_prev, t_prev, t, solver_type="dpm_solver"):   # :889-968
        ns = self.ns
        (m1, m0), (t1, t0) = m_prev, t_prev
        h0 = ns.lam(t0) - ns.lam(t1)
        h = ns.lam(t) - ns.lam(t0)
        r0 = h0 / h
        out = {}
        for k in KEYS:
            d1 = _bc(1.0 / r0, x[k]) * (m0[k] - m1[k])
            if self.predict_x0:
                e = ns.alpha(t) * (torch.exp(-h) - 1.0)
                base = _bc(ns.sigma(t) / ns.sigma(t0), x[k]) * x[k] - _bc(e, x[k]) * m0[k]
                out[k] = (base - 0.5 * _bc(e, x[k]) * d1 if solver_type == "dpm_solver"
                          else base + _bc(ns.alpha(t) * ((torch.exp(-h) - 1.0) / h + 1.0), x[k]) * d1)
            else:
                e = ns.sigma(t) * (torch.exp(h) - 1.0)
                base = _bc(torch.exp(ns.log_mean(t) - ns.log_mean(t0)), x[k]) * x[k] - _bc(e, x[k]) * m0[k]
                out[k] = (base - 0.5 * _bc(e, x[k]) * d1 if solver_type == "dpm_solver"
                          else base - _bc(ns.sigma(t) * ((torch.exp(h) - 1.0) / h - 1.0), x[k]) * d1)
        return out

    def multi3(self, x, m_prev, t_prev, t):             # :970-1036 with the intended (per-sample) audio broadcast
        ns = self.ns
        (m2, m1, m0), (t2, t1, t0) = m_prev, t_prev
        h1, h0, h = ns.lam(t1) - ns.lam(t2), ns.lam(t0) - ns.lam(t1), ns.lam(t) - ns.lam(t0)
        r0, r1 = h0 / h, h1 / h
        out = {}
        for k in KEYS:
            d10 = _bc(1.0 / r0, x[k]) * (m0[k] - m1[k])
            d11 = _bc(1.0 / r1, x[k]) * (m1[k] - m2[k])
            d1 = d10 + _bc(r0 / (r0 + r1), x[k]) * (d10 - d11)
            d2 = _bc(1.0 / (r0 + r1), x[k]) * (d10 - d11)
            if self.predict_x0:
                al = ns.alpha(t)
                out[k] = (_bc(ns.sigma(t) / ns.sigma(t0), x[k]) * x[k] - _bc(al * (torch.exp(-h) - 1.0), x[k]) * m0[k]
                          + _bc(al * ((torch.exp(-h) - 1.0) / h + 1.0), x[k]) * d1
                          - _bc(al * ((torch.exp(-h) - 1.0 + h) / h ** 2 - 0.5), x[k]) * d2)
            else:
                sg = ns.sigma(t)
                out[k] = (_bc(torch.exp(ns.log_mean(t) - ns.log_mean(t0)), x[k]) * x[k]
                          - _bc(sg * (torch.exp(h) - 1.0), x[k]) * m0[k]
                          - _bc(sg * ((torch.exp(h) - 1.0) / h - 1.0), x[k]) * d1
                          - _bc(sg * ((torch.exp(h) - 1.0 - h) / h ** 2 - 0.5), x[k]) * d2)
        return out

    def multi(self, x, m_prev, t_prev, t, order, solver_type):
        if order == 1:
            return self.first(x, t_prev[-1], t, m_s=m_prev[-1])
        if order == 2:
            return self.multi2(x, m_prev, t_prev, t, solver_type)
        return self.multi3(x, m_prev, t_prev, t)

    # ---- drivers
    def time_steps(self, skip_type, t_T, t_0, n):        # :451-478
        if skip_type == "logSNR":
            l_T, l_0 = self.ns.lam(torch.tensor(t_T)), self.ns.lam(torch.tensor(t_0))
            return self.ns.inv_lam(torch.linspace(l_T.item(), l_0.item(), n + 1))
        if skip_type == "time_uniform":
            return torch.linspace(t_T, t_0, n + 1)
        return torch.linspace(t_T ** 0.5, t_0 ** 0.5, n + 1).pow(2)

    @staticmethod
    def orders(steps, order):                            # :480-524
        if order == 3:
            k = steps // 3 + 1
            return [3] * (k - 2) + [2, 1] if steps % 3 == 0 else ([3] * (k - 1) + [1] if steps % 3 == 1 else [3] * (k - 1) + [2])
        if order == 2:
            return [2] * (steps // 2) + ([1] if steps % 2 else [])
        return [1] * steps

    def adaptive(self, x, order, t_T, t_0, h_init=0.05, atol=0.0078, rtol=0.05, theta=0.9, t_err=1e-5,
                 solver_type="dpm_solver"):               # :1088-1149
        ns = self.ns
        b = x["video"].shape[0]
        s = t_T * torch.ones(b)
        lam_s, lam_0 = ns.lam(s), ns.lam(t_0 * torch.ones(b))
        h = h_init * torch.ones(b)
        x_prev = x
        while torch.abs(s - t_0).mean() > t_err:
            t = ns.inv_lam(lam_s + h)
            if order == 2:
                lo, kw = self.first(x, s, t, ret=True)
                hi = self.second(x, s, t, r1=0.5, m_s=kw["m_s"], solver_type=solver_type)
            else:
                lo, kw = self.second(x, s, t, r1=1.0 / 3.0, ret=True, solver_type=solver_type)
                hi = self.third(x, s, t, r1=1.0 / 3.0, r2=2.0 / 3.0, m_s=kw["m_s"], m_s1=kw["m_s1"])
            errs = []
            for k in KEYS:
                delta = torch.max(torch.full_like(x[k], atol), rtol * torch.max(lo[k].abs(), x_prev[k].abs()))
                e = ((hi[k] - lo[k]) / delta).reshape(b, -1)
                errs.append(torch.sqrt((e * e).mean(dim=-1, keepdim=True)))
            E = torch.cat(errs).max()
            if torch.all(E <= 1.0):
                x, s, x_prev = hi, t, lo
                lam_s = ns.lam(s)
            h = torch.min(theta * h * torch.float_power(E, -1.0 / order).float(), lam_0 - lam_s)
        return x

    def sample(self, x, steps=20, order=3, skip_type="time_uniform", method="singlestep", denoise=False,
               solver_type="dpm_solver", atol=0.0078, rtol=0.05):   # :1151-1298
        ns = self.ns
        t_0, t_T = 1.0 / ns.N, ns.T
        b = x["video"].shape[0]
        self.model_times = []
        with torch.no_grad():
            if method == "adaptive":
                x = self.adaptive(x, order, t_T, t_0, atol=atol, rtol=rtol, solver_type=solver_type)
            elif method == "multistep":
                ts = self.time_steps(skip_type, t_T, t_0, steps)
                vt = ts[0].expand(b)
                m_prev, t_prev = [self.f(x, vt)], [vt]
                for o in range(1, order):
                    vt = ts[o].expand(b)
                    x = self.multi(x, m_prev, t_prev, vt, o, solver_type)
                    m_prev.append(self.f(x, vt))
                    t_prev.append(vt)
                for step in range(order, steps + 1):
                    vt = ts[step].expand(b)
                    x = self.multi(x, m_prev, t_prev, vt, order, solver_type)
                    m_prev, t_prev = m_prev[1:] + [None], t_prev[1:] + [vt]
                    if step < steps:
                        m_prev[-1] = self.f(x, vt)
            else:
                if method == "singlestep":
                    orders = self.orders(steps, order)
                    ts = self.time_steps(skip_type, t_T, t_0, steps)
                else:
                    orders = [order] * (steps // order)
                    ts = self.time_steps(skip_type, t_T, t_0, (steps // order) * order)
                i = 0
                for o in orders:
                    vs, vt = ts[i].expand(b), ts[i + o].expand(b)
                    h = ns.lam(ts[i + o]) - ns.lam(ts[i])
                    r1 = None if o <= 1 else (ns.lam(ts[i + 1]) - ns.lam(ts[i])) / h
                    r2 = None if o <= 2 else (ns.lam(ts[i + 2]) - ns.lam(ts[i])) / h
                    if o == 1:
                        x = self.first(x, vs, vt)
                    elif o == 2:
                        x = self.second(x, vs, vt, r1=r1, solver_type=solver_type)
                    else:
                        x = self.third(x, vs, vt, r1=r1, r2=r2)
                    i += o
            if denoise:
                x = self.x0(x, torch.ones(b) * t_0)
        return x
