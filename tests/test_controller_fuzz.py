"""Randomised parity of the C-ABI controller (`mc_ctrl_*`, one parameterisation per adapter: magcache_b200.config.FAMILIES) against
the independent pure-Python restatement oracle/controller_ref.py::AdapterControllerRef, which is first pinned against the golden
schedules produced by the reference's own statements. Bit-exact: skip mask, counter and float64 accumulators, on random tables,
thresholds, K, retention ratios, step counts and expert boundaries (hypothesis)."""
import ctypes
import json
import os

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as hs

from magcache_b200 import _lib as L
from magcache_b200.config import FAMILIES, INITIAL_ACCUMULATED_STEPS, MagCacheConfig, interp_cfg, nearest_interp
from magcache_b200.controller import make_ctrl_config
from oracle.controller_ref import AdapterControllerRef

G = os.path.join(os.path.dirname(__file__), "golden")
with open(os.path.join(G, "masks.json")) as f:
    MASKS = json.load(f)
with open(os.path.join(G, "extra_adapters.json")) as f:
    EXTRA = json.load(f)
with open(os.path.join(G, "paper_eval_adapters.json")) as f:
    PAPER = json.load(f)
with open(os.path.join(G, "tables_all.json")) as f:
    TABLES = json.load(f)


# ---------------------------------------------------------------------------------------------- pin the Python restatement
def test_restatement_pinned_wan_flux_hunyuan():
    for c in MASKS:
        t = np.array(TABLES[c["table"]]["values"])
        if c["family"] == "wan2.1":
            n, ratios = 2 * c["steps"], interp_cfg(t, c["steps"])
        else:
            n = c["steps"]
            ratios = t if len(t) == n else nearest_interp(t, n)
        ref = AdapterControllerRef(c["family"], ratios, n, c["thresh"], c["K"], c["R"])
        assert "".join(map(str, ref.mask(c["calls"]))) == c["mask"], (c["family"], c["table"], c["steps"])


def test_restatement_pinned_wan22_omnigen2():
    for c in EXTRA["wan22_masks"]:
        ratios = interp_cfg(np.array(EXTRA["tables"][c["table"]]["values"]), c["steps"])
        fam = {"t2v": "wan2.2-t2v", "i2v": "wan2.2-i2v", "ti2v": "wan2.2-ti2v"}[c["mode"]]
        split = None if c["high_noise_steps"] is None else 2 * c["high_noise_steps"]
        ref = AdapterControllerRef(fam, ratios, 2 * c["steps"], c["thresh"], c["K"], c["R"], split_step=split)
        assert "".join(map(str, ref.mask(c["calls"]))) == c["mask"]
        assert ref.err == c["final"]["accumulated_err"] and ref.ratio == c["final"]["accumulated_ratio"]
    for c in EXTRA["omnigen2_masks"]:
        t = np.array(EXTRA["tables"][c["table"]]["values"])
        ratios = t if len(t) == c["steps"] else nearest_interp(t, c["steps"])
        ref = AdapterControllerRef("omnigen2", ratios, c["steps"], c["thresh"], c["K"], c["R"])
        got = []
        for k in range(c["steps"]):
            ref.cnt = k
            got.append(int(ref.step()))
        assert "".join(map(str, got)) == c["mask"]


def test_restatement_pinned_framepack_eval_opensora():
    for c in PAPER["framepack_masks"]:
        ref = AdapterControllerRef("framepack", c["ratios"], c["steps"], c["thresh"], c["K"], c["R"])
        assert "".join(map(str, ref.mask(c["calls"]))) == c["mask"]
    for c in PAPER["eval_wan_masks"]:
        ref = AdapterControllerRef("wan2.1-eval", PAPER["tables"]["wan2.1_eval"]["values"], 2 * c["steps"], c["thresh"], c["K"])
        assert "".join(map(str, ref.mask(c["calls"]))) == c["mask"]
    for c in PAPER["opensora_masks"]:
        ref = AdapterControllerRef("opensora", PAPER["tables"]["opensora_eval"]["values"], 30, c["thresh"], c["K"], skip_time=c["skip_time"])
        assert "".join(map(str, ref.mask(c["calls"]))) == c["mask"]
        assert ref.err[0] == c["final"]["accumulated_err"]


# ---------------------------------------------------------------------------------------------- fuzz: C ABI == restatement
def _run_c(family, ratios, n, thresh, K, R, calls, split=None, skip_time=None):
    kw = dict(FAMILIES[family])
    if family in ("wan2.2-t2v", "wan2.2-i2v"):
        kw["split_step"] = split
    if family == "opensora":
        kw["split_step"] = skip_time
    if family == "wan2.1-eval":
        R = 0.2  # hard-coded in that forward (wan_magcache.py:772)
    cfg = make_ctrl_config(n, thresh, K, R, ratios, **kw)
    st = L.CtrlState()
    st.accumulated_ratio[0] = st.accumulated_ratio[1] = 1.0
    st.accumulated_steps[0] = INITIAL_ACCUMULATED_STEPS.get(family, 0)
    skip = ctypes.c_int32()
    mask = []
    for _ in range(calls):
        rc = L.lib.mc_ctrl_decide(ctypes.byref(cfg), ctypes.byref(st), ctypes.byref(skip))
        if rc != 0:
            return None, None
        mask.append(skip.value)
        L.check(L.lib.mc_ctrl_advance(ctypes.byref(cfg), ctypes.byref(st)))
    return mask, st


FUZZ_FAMILIES = [f for f in FAMILIES if f != "omnigen2"]  # omnigen2's cnt is driven by its sampler: covered by the golden cases


@settings(max_examples=400, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(family=hs.sampled_from(FUZZ_FAMILIES), steps=hs.integers(4, 60), seed=hs.integers(0, 2 ** 31 - 1),
       thresh=hs.sampled_from([0.005, 0.02, 0.06, 0.12, 0.24, 0.5, 2.0]), K=hs.integers(1, 8),
       R=hs.sampled_from([0.0, 0.05, 0.1, 0.2, 0.25, 1.0 / 3.0, 0.4, 0.57, 0.7]), spread=hs.sampled_from([0.002, 0.01, 0.05, 0.2]),
       split_frac=hs.floats(0.1, 0.9), skip_time=hs.integers(1, 12))
def test_c_abi_equals_restatement_on_random_configurations(family, steps, seed, thresh, K, R, spread, split_frac, skip_time):
    rng = np.random.default_rng(seed)
    branches = FAMILIES[family]["branches"]
    n = steps * branches
    if family == "opensora":
        n = max(n, skip_time + 2)
    off = FAMILIES[family].get("table_offset", 0)
    if family == "wan2.1-eval":
        n = max(n, 50)  # upstream's ratio[t-10] needs int(n*0.2) >= 10
    table = np.concatenate([np.ones(branches), 1.0 + spread * rng.standard_normal(n)])[: n - off if off else n]
    if off:
        table = 1.0 + spread * rng.standard_normal(n - off)
    split = None
    if family in ("wan2.2-t2v", "wan2.2-i2v"):
        split = 2 * max(1, int(steps * split_frac))
        split = min(split, n)
    calls = 2 * n + 3
    ref = AdapterControllerRef(family, table, n, thresh, K, R, split_step=split, skip_time=skip_time)
    try:
        want = ref.mask(calls)
    except IndexError:
        want = None
    got, st = _run_c(family, table, n, thresh, K, R, calls, split=split, skip_time=skip_time)
    if want is None:
        assert got is None  # both refuse a table offset that precedes the retention window
        return
    assert got == want, (family, n, thresh, K, R, split)
    assert st.cnt == ref.cnt
    nb = 2 if branches == 2 else 1
    if not (family == "framepack" and ref.cnt == 0):  # FramePack re-initialises lazily at the next call, the C ABI at the wrap
        assert [st.accumulated_err[i] for i in range(nb)] == ref.err
        assert [st.accumulated_ratio[i] for i in range(nb)] == ref.ratio
        assert [st.accumulated_steps[i] for i in range(nb)] == ref.steps


def test_magcacheconfig_schedule_equals_restatement_for_every_preset():
    from magcache_b200 import PRESETS
    for name, cfg in PRESETS.items():
        kw = cfg.ctrl_kwargs()
        ref = AdapterControllerRef(cfg.family, cfg.resolved_ratios(), cfg.num_steps, cfg.thresh, cfg.K, cfg.retention_ratio,
                                   split_step=kw.get("split_step") if cfg.family.startswith("wan2.2") else None,
                                   skip_time=kw.get("split_step") if cfg.family == "opensora" else None)
        if cfg.family == "omnigen2":
            want = []
            for k in range(cfg.num_steps):
                ref.cnt = k
                want.append(int(ref.step()))
        else:
            want = ref.mask(cfg.num_steps)
        assert cfg.schedule().tolist() == want, name
    assert isinstance(MagCacheConfig("hunyuan", table="hunyuan_544p").schedule(), np.ndarray)


# ---------------------------------------------------------------------------------------------- nearest_interp / TeaCache fuzz
@settings(max_examples=300, deadline=None)
@given(L_=hs.integers(1, 200), T_=hs.integers(1, 200), seed=hs.integers(0, 2 ** 31 - 1))
def test_nearest_interp_equals_numpy_expression(L_, T_, seed):
    """`np.round(np.arange(T) * (L-1)/(T-1)).astype(int)` (round half to even) for every (L, T), incl. the linspace form of Qwen-Image."""
    from oracle.controller_ref import nearest_interp as ref_interp
    src = np.random.default_rng(seed).standard_normal(L_)
    assert nearest_interp(src, T_).tolist() == ref_interp(src, T_).tolist()
    out = np.empty(T_)
    dp = ctypes.POINTER(ctypes.c_double)
    L.check(L.lib.mc_nearest_interp_linspace(np.ascontiguousarray(src).ctypes.data_as(dp), L_, out.ctypes.data_as(dp), T_))
    want = src if L_ == T_ else src[np.round(np.linspace(0, L_ - 1, T_)).astype(int)]  # MagCache4QwenImage/magcache_generate.py:14-21
    assert out.tolist() == np.asarray(want).tolist()


@settings(max_examples=300, deadline=None)
@given(steps=hs.integers(3, 40), seed=hs.integers(0, 2 ** 31 - 1), thresh=hs.sampled_from([0.02, 0.08, 0.2, 0.5]), ret=hs.integers(0, 10),
       cut=hs.integers(0, 6), ncoef=hs.integers(1, 8))
def test_teacache_controller_equals_numpy_poly1d_restatement(steps, seed, thresh, ret, cut, ncoef):
    """wan_teacache.py:535-564 restated with np.poly1d on random coefficients / distances."""
    rng = np.random.default_rng(seed)
    n = 2 * steps
    coef = (rng.standard_normal(ncoef) * np.logspace(ncoef - 1, 0, ncoef)).tolist()
    cfg = L.TeaConfig()
    cfg.num_steps, cfg.ret_steps, cfg.cutoff_steps, cfg.n_coef, cfg.thresh = n, ret, n - cut, ncoef, thresh
    for i, c in enumerate(coef):
        cfg.coef[i] = c
    st = L.TeaState()
    calc = ctypes.c_int32()
    acc, cnt = [0, 0], 0
    poly = np.poly1d(coef)
    for _ in range(2 * n + 3):
        d = float(abs(rng.standard_normal()) * 0.1)
        i = cnt % 2
        if cnt < ret or cnt >= n - cut:
            want, acc[i] = True, 0
        else:
            acc[i] += poly(d)
            if acc[i] < thresh:
                want = False
            else:
                want, acc[i] = True, 0
        L.check(L.lib.mc_tea_decide(ctypes.byref(cfg), ctypes.byref(st), d, ctypes.byref(calc)))
        assert bool(calc.value) == want
        assert [st.accumulated[0], st.accumulated[1]] == [float(acc[0]), float(acc[1])]
        L.check(L.lib.mc_tea_advance(ctypes.byref(cfg), ctypes.byref(st)))
        cnt = 0 if cnt + 1 >= n else cnt + 1
        assert st.cnt == cnt
