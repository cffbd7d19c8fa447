"""The caller loop's per-step tail — CFG combine + scheduler update — on the fused `mc_cfg_step` kernel.

Reference call site: eval/magcache/experiments/Wan2.1_EVAL/wan_magcache.py:289-310

    noise_pred = noise_pred_uncond + guide_scale * (noise_pred_cond - noise_pred_uncond)
    temp_x0 = sample_scheduler.step(noise_pred.unsqueeze(0), t, latents[0].unsqueeze(0), return_dict=False, generator=seed_g)[0]

The schedulers (`FlowUniPCMultistepScheduler`, `FlowDPMSolverMultistepScheduler`, :263-279) are upstream Wan code that is not in
the reference tree ("parity unpinned", see oracle/sampler_ref.py). All of them update the latent by a LINEAR combination of the
current sample, the guided flow prediction and a few stored tensors, with scalar coefficients that depend only on the sigma
schedule — so the host computes the scalars in float64 and every tensor operation of a step is one or two launches of
`mc_cfg_step` (instead of ~15 element-wise torch kernels and their temporaries).
"""
import math

import torch

from . import ops


def sampling_sigmas(steps, shift):
    """`shift*s/(1+(shift-1)*s)` over s = linspace(1, 0, steps+1) (float64, terminal 0 included); timestep = 1000*sigma."""
    out = []
    for i in range(steps + 1):
        s = 1.0 - i / steps
        out.append(shift * s / (1.0 + (shift - 1.0) * s))
    return out


def cfg_denoise_step(model, x, t, context, context_null, seq_len, guide_scale, coef_v, coef_x=1.0, forward=None, **kw):
    """`x <- coef_x * x + coef_v * (uncond + g * (cond - uncond))` with both predictions from the patched forward of `model`
    (SURVEY §8f rank 1 as written). On the unsharded Wan engine the combine and the update ride in the epilogue of the unconditional
    call's head kernel (`mc_head_unpatchify_step`): the unconditional prediction is never written, no separate pass over the
    latents runs — on a step where both calls hit the cache the whole step is two streaming head launches. Token-sharded engines
    (whose head stores go to every peer) and multistep solvers that keep history terms use `ops.cfg_step` after the two calls.
    Bit-equal either way. `forward(x, t, context) -> prediction` replaces the plain `model([x], t=..., context=[...])` call (a
    caller that wraps its forwards, e.g. with timing events)."""
    from .patch import _engine
    if forward is None:
        def forward(xx, tt, cc):
            return model([xx], t=tt, context=[cc], seq_len=seq_len, **kw)[0]
    cond = forward(x, t, context)
    eng = _engine(model)
    # the fused form is built for one timestep per call and 16 output channels (not Wan2.2 TI2V-5B: per-token t, 48 channels)
    if getattr(eng, "shard", None) is None and hasattr(eng, "arm_step") and t.numel() == 1 and len(getattr(eng, "head_groups", (0,))) == 1:
        eng.arm_step(cond, x, guide_scale, coef_x, coef_v, out=x)
        return forward(x, t, context_null)
    uncond = forward(x, t, context_null)
    return ops.cfg_step(cond, uncond, guide_scale, x, coef_v, coef_x=coef_x, out=x)


class FlowEulerSampler:
    """x <- x + (sigma_{i+1} - sigma_i) * v : one launch per step, in place."""

    def __init__(self, sigmas):
        self.sigmas = [float(s) for s in sigmas]
        self.i = 0

    @property
    def timestep(self):
        return 1000.0 * self.sigmas[self.i]

    def step(self, cond, uncond, guide_scale, x):
        d = self.sigmas[self.i + 1] - self.sigmas[self.i]
        self.i += 1
        return ops.cfg_step(cond, uncond, guide_scale, x, d, out=x)

    def denoise(self, model, x, t, context, context_null, seq_len, guide_scale, **kw):
        """One whole step of the caller loop (wan_magcache.py:296-310): cond call, uncond call, CFG combine, scheduler update; the
        latent `x` [C, F, H, W] is updated in place and returned. `model` carries the patched forward (`init_magcache`)."""
        d = self.sigmas[self.i + 1] - self.sigmas[self.i]
        self.i += 1
        return cfg_denoise_step(model, x, t, context, context_null, seq_len, guide_scale, coef_v=d, **kw)


class FlowUniPCSampler:
    """UniPC (bh2, x0-prediction, order 2, lower order on the first and last step, corrector after the first step) on the
    flow parameterisation alpha = 1 - sigma. Two launches per step: [corrector + x0-prediction] and [predictor]."""

    def __init__(self, sigmas, order=2):
        if order not in (1, 2):
            raise NotImplementedError("FlowUniPCSampler: solver order 1 or 2 (the reference's callers use the default, 2)")
        self.sigmas = [float(s) for s in sigmas]
        self.order = order
        self.i = 0
        self.lower_order_nums = 0
        self.this_order = 1
        self._m = []       # x0 predictions, newest last (device tensors, recycled)
        self._last = None  # corrected sample of the previous step
        self._free = []

    @property
    def timestep(self):
        return 1000.0 * self.sigmas[self.i]

    @staticmethod
    def _lam(sigma):
        return math.log(1.0 - sigma) - math.log(sigma)

    @staticmethod
    def _b(hh, order):
        """b_k = h_phi_{k+1} * k! / B(h) of the UniPC linear system (k = 1..order), B(h) = expm1(hh)."""
        h_phi_1 = math.expm1(hh)
        B_h = h_phi_1
        h_phi_k = h_phi_1 / hh - 1.0
        b, fact = [], 1
        for k in range(1, order + 1):
            b.append(h_phi_k * fact / B_h)
            fact *= k + 1
            h_phi_k = h_phi_k / hh - 1.0 / fact
        return h_phi_1, B_h, b

    def _buf(self, like):
        return self._free.pop() if self._free else torch.empty_like(like)

    def corrector_coeffs(self):
        """Coefficients of x_corr = c_last*last + c_m0*m0 + c_m1*m1 + c_mt*m_t at step i (order = self.this_order)."""
        s_t, s_0 = self.sigmas[self.i], self.sigmas[self.i - 1]
        alpha_t = 1.0 - s_t
        h = self._lam(s_t) - self._lam(s_0)
        h_phi_1, B_h, b = self._b(-h, self.this_order)
        if self.this_order == 1:
            rho_t, c_m1, extra_m0 = 0.5, 0.0, 0.0
        else:
            rk = (self._lam(self.sigmas[self.i - 2]) - self._lam(s_0)) / h
            # [[1, 1], [rk, 1]] @ [rho0, rho_t] = [b1, b2]
            rho0 = (b[0] - b[1]) / (1.0 - rk)
            rho_t = b[0] - rho0
            c_m1 = -alpha_t * B_h * rho0 / rk
            extra_m0 = alpha_t * B_h * rho0 / rk
        c_m0 = -alpha_t * h_phi_1 + alpha_t * B_h * rho_t + extra_m0
        return s_t / s_0, c_m0, c_m1, -alpha_t * B_h * rho_t

    def predictor_coeffs(self, order):
        """Coefficients of x_next = c_x*x_corr + c_mt*m_t + c_mp*m_prev at step i."""
        s_t, s_0 = self.sigmas[self.i + 1], self.sigmas[self.i]
        if s_t == 0.0:  # terminal sigma: lambda = +inf, expm1(-inf) = -1 -> the sample becomes the x0 prediction
            return 0.0, 1.0, 0.0
        alpha_t = 1.0 - s_t
        h = self._lam(s_t) - self._lam(s_0)
        h_phi_1, B_h, _ = self._b(-h, order)
        c_mt, c_mp = -alpha_t * h_phi_1, 0.0
        if order == 2:
            rk = (self._lam(self.sigmas[self.i - 1]) - self._lam(s_0)) / h
            c_mp = -alpha_t * B_h * 0.5 / rk
            c_mt -= c_mp
        return s_t / s_0, c_mt, c_mp

    def step(self, cond, uncond, guide_scale, x):
        """cond / uncond: the two forwards' outputs at (x, sigma_i), fp32, same shape as x. Returns the next sample (a tensor
        owned by the sampler until the next call)."""
        sig = self.sigmas[self.i]
        m_t, x_corr = self._buf(x), self._buf(x)
        if self.i > 0 and self._last is not None:
            c_last, c_m0, c_m1, c_mt = self.corrector_coeffs()
            hist, coef = [self._last, self._m[-1]], [c_last, c_m0]
            if self.this_order == 2:
                hist.append(self._m[-2])
                coef.append(c_m1)
            # m_t = x - sigma*v, so c_mt*m_t = c_mt*x - c_mt*sigma*v
            ops.cfg_step(cond, uncond, guide_scale, x, -c_mt * sig, coef_x=c_mt, hist=hist, coef_h=coef, sigma=sig, out=x_corr, x0_out=m_t)
        else:
            ops.cfg_step(cond, uncond, guide_scale, x, 0.0, coef_x=1.0, sigma=sig, out=x_corr, x0_out=m_t)
        self._m.append(m_t)
        if len(self._m) > self.order:
            self._free.append(self._m.pop(0))
        n_steps = len(self.sigmas) - 1
        self.this_order = min(min(self.order, n_steps - self.i), self.lower_order_nums + 1)
        if self._last is not None:
            self._free.append(self._last)
        self._last = x_corr
        c_x, c_mt, c_mp = self.predictor_coeffs(self.this_order)
        out = self._buf(x)
        hist, coef = ([self._m[-2]], [c_mp]) if self.this_order == 2 else ([], [])
        # second launch: no guidance (cond = uncond = m_t, g = 0 gives v = m_t exactly)
        ops.cfg_step(m_t, m_t, 0.0, x_corr, c_mt, coef_x=c_x, hist=hist, coef_h=coef, out=out)
        if self.lower_order_nums < self.order:
            self.lower_order_nums += 1
        self.i += 1
        return out
