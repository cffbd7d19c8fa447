"""ORACLE — test infrastructure only. numpy/Python restatement of the MagCache host logic, one function per reference site.

Pinned against tests/golden/{nearest_interp,masks}.json, which were produced by executing the reference's own statements
(tests/golden/make_golden.py).
"""
import math

import numpy as np


def nearest_interp(src_array, target_length):
    """MagCache4Wan2.1/magcache_generate.py:27-34 (same body in MagCache4FLUX/magcache_flux.py:12-19 and
    MagCache4HunyuanVideo/magcache_sample_video.py:20-27)."""
    src_array = np.asarray(src_array)
    n = len(src_array)
    if target_length == 1:
        return np.array([src_array[-1]])
    step = (n - 1) / (target_length - 1)
    picks = np.round(np.arange(target_length) * step).astype(int)  # np.round: half to even
    return src_array[picks]


def interp_cfg(table, sample_steps):
    """Per-branch interpolation + re-interleave, MagCache4Wan2.1/magcache_generate.py:915-919."""
    table = np.asarray(table, dtype=np.float64)
    if len(table) == sample_steps * 2:
        return table
    cond = nearest_interp(table[0::2], sample_steps)
    uncond = nearest_interp(table[1::2], sample_steps)
    return np.stack([cond, uncond], axis=1).reshape(-1)


class ControllerRef:
    """State machine of the skip decision. family in {"wan2.1", "flux", "hunyuan"}:

    wan2.1  : per-CFG-branch lists, `<`,  start int(n*R)          magcache_generate.py:277-292, :306-311
    flux    : scalars, `<=`, start int(R*n+0.5), step-11 veto      magcache_flux.py:326-338, :431-436
    hunyuan : scalars, `<=`, start int(R*n)                        magcache_sample_video.py:88-102, :149-154
    """

    def __init__(self, family, mag_ratios, num_steps, thresh, K, retention_ratio):
        self.family, self.mag_ratios = family, np.asarray(mag_ratios, dtype=np.float64)
        self.num_steps, self.thresh, self.K, self.R = num_steps, thresh, K, retention_ratio
        self.cnt = 0
        self._reset()

    def _reset(self):
        nb = 2 if self.family == "wan2.1" else 1
        self.ratio, self.err, self.steps = [1.0] * nb, [0.0] * nb, [0] * nb

    def _start(self):
        if self.family == "flux":
            return int(self.R * self.num_steps + 0.5)
        return int(self.num_steps * self.R)

    def step(self):
        """One forward call: returns True when the transformer stack is skipped; advances the counter."""
        skip = False
        if self.cnt >= self._start():
            i = self.cnt % 2 if self.family == "wan2.1" else 0
            self.ratio[i] = self.ratio[i] * self.mag_ratios[self.cnt]
            self.steps[i] += 1
            self.err[i] += np.abs(1 - self.ratio[i])
            if self.family == "wan2.1":
                ok = self.err[i] < self.thresh and self.steps[i] <= self.K
            else:
                ok = self.err[i] <= self.thresh and self.steps[i] <= self.K
            if self.family == "flux":
                ok = ok and np.round(self.cnt * ((28 - 1) / (self.num_steps - 1))).astype(int) != 11
            if ok:
                skip = True
            else:
                self.ratio[i], self.steps[i], self.err[i] = 1.0, 0, 0.0
        self.cnt += 1
        if self.cnt >= self.num_steps:
            self.cnt = 0
            self._reset()
        return skip

    def mask(self, calls):
        return [1 if self.step() else 0 for _ in range(calls)]


def calibration_stats(residual, previous, denom_eps=0.0):
    """MagCache4Wan2.1/magcache_generate.py:167-169 on torch tensors (eval variant adds 1e-8: wan_magcache.py:652-654)."""
    import torch.nn.functional as F
    ratio = residual.norm(dim=-1) / (previous.norm(dim=-1) + denom_eps)
    return ratio.mean().item(), ratio.std().item(), (1 - F.cosine_similarity(residual, previous, dim=-1, eps=1e-8)).mean().item()
