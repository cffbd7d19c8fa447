"""ORACLE — test infrastructure only (imported by tests/, never by the product).

Tensor-form restatement of the caller loop's sampler step (eval/magcache/experiments/Wan2.1_EVAL/wan_magcache.py:263-310):
classifier-free-guidance combine followed by `sample_scheduler.step(...)`.

PARITY UNPINNED: the schedulers themselves (`FlowUniPCMultistepScheduler`, `FlowDPMSolverMultistepScheduler`,
`get_sampling_sigmas`) live in upstream Wan-Video/Wan2.1 `wan/utils/fm_solvers*.py`, which is NOT under /root/reference and
is not pinned by the reference (README: "clone the repo"). What follows restates the published UniPC algorithm
(Zhao et al. 2023, B(h) = expm1(h) "bh2" variant, data/x0-prediction form, order 2, lower order at the first and last step)
on the flow-matching parameterisation alpha_t = 1 - sigma_t the Wan schedulers use; the reference's own call site fixes only
the call order (cond, uncond, combine, step) and the solver names.
"""

import torch


def sampling_sigmas(steps, shift):
    """sigma_i = shift*s/(1 + (shift-1)*s) for s = linspace(1, 0, steps+1); the trailing 0 is the terminal sigma."""
    s = torch.linspace(1.0, 0.0, steps + 1, dtype=torch.float64)
    return shift * s / (1.0 + (shift - 1.0) * s)


def cfg(cond, uncond, guide):
    return uncond + guide * (cond - uncond)  # wan_magcache.py:301-302


class EulerRef:
    def __init__(self, sigmas):
        self.sigmas, self.i = sigmas, 0

    def step(self, v, x):
        out = x + (self.sigmas[self.i + 1] - self.sigmas[self.i]) * v
        self.i += 1
        return out


class UniPCRef:
    """UniPC-bh2, predict-x0, solver_order 2, lower_order_final, corrector enabled on every step after the first."""

    def __init__(self, sigmas, order=2):
        self.sigmas = sigmas.double()
        self.order = order
        self.i = 0
        self.model_outputs = [None] * order
        self.lower_order_nums = 0
        self.last_sample = None
        self.this_order = 1

    @staticmethod
    def _lam(sigma):
        alpha = 1.0 - sigma
        return torch.log(alpha) - torch.log(sigma)

    def _coeffs(self, h, rks, order, corrector):
        hh = -h
        h_phi_1 = torch.expm1(hh)
        h_phi_k = h_phi_1 / hh - 1.0
        B_h = torch.expm1(hh)
        R, b, fact = [], [], 1
        rks = torch.stack(rks)
        for k in range(1, order + 1):
            R.append(rks ** (k - 1))
            b.append(h_phi_k * fact / B_h)
            fact *= k + 1
            h_phi_k = h_phi_k / hh - 1.0 / fact
        R, b = torch.stack(R), torch.stack(b)
        return h_phi_1, B_h, R, b

    def _predict(self, x, order):
        m0 = self.model_outputs[-1]
        s_t, s_0 = self.sigmas[self.i + 1], self.sigmas[self.i]
        alpha_t = 1.0 - s_t
        if float(s_t) == 0.0:  # terminal step: lambda_t = +inf, expm1(-inf) = -1  ->  x_t = m0
            return m0.clone()
        h = self._lam(s_t) - self._lam(s_0)
        rks, D1s = [], []
        for k in range(1, order):
            mk = self.model_outputs[-(k + 1)]
            rk = (self._lam(self.sigmas[self.i - k]) - self._lam(s_0)) / h
            rks.append(rk)
            D1s.append((mk - m0) / rk)
        rks.append(torch.tensor(1.0, dtype=torch.float64))
        h_phi_1, B_h, R, b = self._coeffs(h, rks, order, False)
        x_t = s_t / s_0 * x - alpha_t * h_phi_1 * m0
        if D1s:
            rhos = torch.tensor([0.5], dtype=torch.float64) if order == 2 else torch.linalg.solve(R[:-1, :-1], b[:-1])
            x_t = x_t - alpha_t * B_h * sum(r * d for r, d in zip(rhos, D1s))
        return x_t

    def _correct(self, m_t, last_sample, order):
        m0 = self.model_outputs[-1]
        s_t, s_0 = self.sigmas[self.i], self.sigmas[self.i - 1]
        alpha_t = 1.0 - s_t
        h = self._lam(s_t) - self._lam(s_0)
        rks, D1s = [], []
        for k in range(1, order):
            mk = self.model_outputs[-(k + 1)]
            rk = (self._lam(self.sigmas[self.i - (k + 1)]) - self._lam(s_0)) / h
            rks.append(rk)
            D1s.append((mk - m0) / rk)
        rks.append(torch.tensor(1.0, dtype=torch.float64))
        h_phi_1, B_h, R, b = self._coeffs(h, rks, order, True)
        rhos = torch.tensor([0.5], dtype=torch.float64) if order == 1 else torch.linalg.solve(R, b)
        x_t = s_t / s_0 * last_sample - alpha_t * h_phi_1 * m0
        corr = sum(r * d for r, d in zip(rhos[:-1], D1s)) if D1s else 0.0
        return x_t - alpha_t * B_h * (corr + rhos[-1] * (m_t - m0))

    def step(self, v, x):
        """v: guided model output at (x, sigma_i). Returns the next sample."""
        x, v = x.double(), v.double()
        m_t = x - self.sigmas[self.i] * v  # flow prediction -> x0 prediction
        if self.i > 0 and self.last_sample is not None:
            x = self._correct(m_t, self.last_sample, self.this_order)
        self.model_outputs = self.model_outputs[1:] + [m_t]
        n_steps = len(self.sigmas) - 1
        this_order = min(self.order, n_steps - self.i)  # lower_order_final
        self.this_order = min(this_order, self.lower_order_nums + 1)
        self.last_sample = x
        out = self._predict(x, self.this_order)
        if self.lower_order_nums < self.order:
            self.lower_order_nums += 1
        self.i += 1
        return out


def denoise(model_v, x, sigmas, guide, sampler="unipc"):
    """model_v(x, sigma, branch) -> flow prediction; runs the caller loop (cond first, then uncond) and returns the final sample."""
    s = UniPCRef(sigmas) if sampler == "unipc" else EulerRef(sigmas)
    for i in range(len(sigmas) - 1):
        v = cfg(model_v(x, sigmas[i], 0), model_v(x, sigmas[i], 1), guide)
        x = s.step(v, x)
    return x
