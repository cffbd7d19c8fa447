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


class AdapterControllerRef:
    """Pure-Python restatement of EVERY controller row of SURVEY Appendix A, one `family` per adapter of the reference. Used as the
    second, independent implementation in the randomised tests (tests/test_controller_fuzz.py); itself pinned against the golden
    schedules produced by the reference's statements (tests/golden/{masks,extra_adapters,paper_eval_adapters}.json).

      wan2.1       MagCache4Wan2.1/magcache_generate.py:277-292, :306-311       per branch, `<`,  int(n*R)
      wan2.1-eval  eval/.../Wan2.1_EVAL/wan_magcache.py:770-786, :807-815       per branch, `<=`, int(n*0.2), ratio[t-10]
      wan2.2-t2v   MagCache4Wan2.2/magcache_generate.py:294-317 (mode t2v)      per branch, `<`,  two-expert window, float32 compare
      wan2.2-i2v   idem (mode i2v)                                               per branch, `<`,  int(split+(n-split)*R)
      wan2.2-ti2v  idem (split_step None)                                        per branch, `<`,  int(n*R)
      qwen-image   MagCache4QwenImage/magcache_generate.py:205-219, :243-244     = wan2.1, but the wrap resets only the counter
      flux         MagCache4FLUX/magcache_flux.py:326-338, :431-436              scalar, `<=`, int(R*n+0.5), step-11 veto
      flux-kontext MagCache4FLUX_Kontext/magcache_flux_kontext.py:328-340        = flux
      hunyuan      MagCache4HunyuanVideo/magcache_sample_video.py:88-102         scalar, `<=`, int(R*n)
      framepack    MagCache4FramePack/magcache_demo_gradio.py:252-270, :298-300  scalar, `<=`, int(R*n) and cnt>=1, |1-r|<=0.06, re-init at 0
      omnigen2     MagCache4OmniGen2/magcache/magcache_utils.py:342-354          scalar, `<=`, ceil(R*n); accumulated_steps starts at 3
      opensora     eval/magcache/experiments/opensora.py:297-308, :348-354      scalar, `<=`, t>=skip_time, ratio[t-1], signed error
    """

    PER_BRANCH = ("wan2.1", "wan2.1-eval", "wan2.2-t2v", "wan2.2-i2v", "wan2.2-ti2v", "qwen-image")

    def __init__(self, family, mag_ratios, num_steps, thresh, K, retention_ratio=0.2, split_step=None, skip_time=None):
        self.family, self.table = family, [float(v) for v in mag_ratios]
        self.n, self.thresh, self.K, self.R = int(num_steps), thresh, int(K), retention_ratio
        self.split, self.skip_time = split_step, skip_time
        self.nb = 2 if family in self.PER_BRANCH else 1
        self.cnt = 0
        self.ratio, self.err, self.steps = [1.0] * self.nb, [0.0] * self.nb, [3 if family == "omnigen2" else 0] * self.nb

    def _eligible(self, c):
        f, n, R = self.family, self.n, self.R
        if f == "wan2.1-eval":
            return c >= int(n * 0.2)
        if f == "wan2.2-t2v" and self.split is not None:
            import torch  # upstream keeps cnt in an int64 tensor: the upper bound is compared in float32
            upper = (n - self.split) * R + self.split
            return not (c < int(self.split * R) or (bool(torch.tensor(c) <= upper) and c >= self.split))
        if f == "wan2.2-i2v" and self.split is not None:
            return not c < int(self.split + (n - self.split) * R)
        if f in ("flux", "flux-kontext"):
            return c >= int(R * n + 0.5)
        if f == "framepack":
            return c >= int(R * n) and c >= 1
        if f == "omnigen2":
            return c >= math.ceil(R * n)
        if f == "opensora":
            return c >= self.skip_time
        return c >= int(n * R)

    def step(self):
        f, c = self.family, self.cnt
        if f == "framepack" and c == 0:
            self.ratio, self.err, self.steps = [1.0], [0.0], [0]
        skip = False
        if self._eligible(c):
            i = c % 2 if self.nb == 2 else 0
            off = 10 if f == "wan2.1-eval" else (1 if f == "opensora" else 0)
            if c - off < 0:
                raise IndexError("table offset before the retention window")
            cur = self.table[c - off]
            self.ratio[i] = self.ratio[i] * cur
            self.steps[i] += 1
            self.err[i] += (1 - self.ratio[i]) if f == "opensora" else abs(1 - self.ratio[i])
            strict = f in ("wan2.1", "wan2.2-t2v", "wan2.2-i2v", "wan2.2-ti2v", "qwen-image")
            ok = (self.err[i] < self.thresh) if strict else (self.err[i] <= self.thresh)
            ok = ok and self.steps[i] <= self.K
            if f in ("flux", "flux-kontext"):
                ok = ok and int(np.round(c * ((28 - 1) / (self.n - 1)))) != 11
            if f == "framepack":
                ok = ok and abs(1 - cur) <= 0.06
            if ok:
                skip = True
            else:
                self.ratio[i], self.steps[i], self.err[i] = 1.0, 0, 0.0
        if f != "omnigen2":  # OmniGen2's sampler owns cnt; everyone else counts forward calls and wraps
            self.cnt += 1
            if self.cnt >= self.n:
                self.cnt = 0
                if f not in ("framepack", "qwen-image"):  # those two reset only the counter (magcache_demo_gradio.py:299-300; Qwen :243-244)
                    self.ratio, self.err, self.steps = [1.0] * self.nb, [0.0] * self.nb, [0] * self.nb
        return skip

    def mask(self, calls):
        return [1 if self.step() else 0 for _ in range(calls)]
