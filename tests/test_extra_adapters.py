"""SURVEY §8f rank 2 — controller variants of the other adapters, bit-exact against fixtures produced by executing the
reference's own statements (tests/golden/make_golden.py::extra_cases): Wan2.2's two-expert skip windows
(MagCache4Wan2.2/magcache_generate.py:290-317, cnt kept in an int64 torch tensor), Qwen-Image's linspace nearest_interp
(MagCache4QwenImage/magcache_generate.py:14-21), OmniGen2's ceil retention + dataclass initial state
(MagCache4OmniGen2/magcache/magcache_utils.py:41-45, 342-354)."""
import ctypes
import json
import os

import numpy as np
import pytest

from magcache_b200 import _lib as L
from magcache_b200.config import interp_cfg, nearest_interp
from magcache_b200.controller import make_ctrl_config, schedule_mask

with open(os.path.join(os.path.dirname(__file__), "golden", "extra_adapters.json")) as f:
    X = json.load(f)
with open(os.path.join(os.path.dirname(__file__), "golden", "tables_all.json")) as f:
    REMAINING = json.load(f)  # every table literal of the reference, incl. Qwen-Image(-Edit) with their [1.0]*2 prefix


@pytest.mark.parametrize("case", X["wan22_masks"], ids=lambda c: f"{c['table']}-{c['mode']}-s{c['steps']}-h{c['high_noise_steps']}-E{c['thresh']}K{c['K']}R{c['R']:.3f}")
def test_wan22_windows(case):
    tbl = X["tables"][case["table"]]["values"]
    ratios = interp_cfg(tbl, case["steps"])
    mode = {"t2v": L.MC_RETAIN_WAN22_T2V, "i2v": L.MC_RETAIN_WAN22_I2V, "ti2v": L.MC_RETAIN_FLOOR}[case["mode"]]
    split = 0 if case["high_noise_steps"] is None else 2 * case["high_noise_steps"]
    cfg = make_ctrl_config(2 * case["steps"], case["thresh"], case["K"], case["R"], ratios, branches=2, cmp=L.MC_CMP_LT,
                           retention_mode=mode, split_step=split)
    mask = schedule_mask(cfg, case["calls"])
    assert "".join(map(str, mask.tolist())) == case["mask"]
    # final accumulator state through the step-wise API
    st = L.CtrlState()
    st.accumulated_ratio[0] = st.accumulated_ratio[1] = 1.0
    skip = ctypes.c_int32()
    for _ in range(case["calls"]):
        L.check(L.lib.mc_ctrl_decide(ctypes.byref(cfg), ctypes.byref(st), ctypes.byref(skip)))
        L.check(L.lib.mc_ctrl_advance(ctypes.byref(cfg), ctypes.byref(st)))
    assert st.cnt == case["final"]["cnt"]
    assert [st.accumulated_err[i] for i in range(2)] == case["final"]["accumulated_err"]
    assert [st.accumulated_ratio[i] for i in range(2)] == case["final"]["accumulated_ratio"]


def test_wan22_float32_threshold_quirk_is_reproduced():
    """cnt is an int64 tensor upstream, so `cnt <= (n-split)*R + split` is evaluated in float32: a double threshold a hair
    below an integer still admits that integer into the skip-disabled window. Find such a configuration and check the C
    controller follows torch's float32 comparison, not exact arithmetic."""
    import torch
    found = None
    for n in range(40, 301, 2):
        for sp in range(2, n, 2):
            for R in (0.7, 0.58, 0.57, 0.29, 0.35, 0.1, 0.2, 0.3, 1.0 / 3.0):
                up = (n - sp) * R + sp
                c = int(round(up))
                if c > up and bool(torch.tensor(c) <= up) and sp <= c < n:  # exact: outside the window; float32: inside
                    found = (n, sp, R, c)
                    break
            if found:
                break
        if found:
            break
    assert found, "no float32/double disagreement found in the searched range"
    n, sp, R, c = found
    ratios = np.full(n, 0.9999)  # would always skip if the controller were consulted
    cfg = make_ctrl_config(n, 0.5, 100, R, ratios, branches=2, cmp=L.MC_CMP_LT, retention_mode=L.MC_RETAIN_WAN22_T2V, split_step=sp)
    mask = schedule_mask(cfg, n)
    assert mask[c] == 0 and mask[min(c + 2, n - 1)] == 1, (found, mask.tolist())


@pytest.mark.parametrize("case", X["qwen_interp"], ids=lambda c: f"{c['L']}to{c['T']}")
def test_qwen_linspace_interp(case):
    src = np.array(case["src"])
    out = np.empty(case["T"])
    dp = ctypes.POINTER(ctypes.c_double)
    L.check(L.lib.mc_nearest_interp_linspace(src.ctypes.data_as(dp), len(src), out.ctypes.data_as(dp), case["T"]))
    assert out.tolist() == case["out"]


@pytest.mark.parametrize("case", X["omnigen2_masks"], ids=lambda c: f"{c['table']}-s{c['steps']}-E{c['thresh']}K{c['K']}R{c['R']}")
def test_omnigen2_ceil_retention_and_initial_state(case):
    tbl = np.array(X["tables"][case["table"]]["values"])
    ratios = tbl if len(tbl) == case["steps"] else nearest_interp(tbl, case["steps"])
    cfg = make_ctrl_config(case["steps"], case["thresh"], case["K"], case["R"], ratios, branches=1, cmp=L.MC_CMP_LE,
                           retention_mode=L.MC_RETAIN_CEIL)
    st = L.CtrlState()
    st.accumulated_ratio[0] = st.accumulated_ratio[1] = 1.0
    st.accumulated_steps[0] = case["initial_accumulated_steps"]  # MagCacheParams default, magcache_utils.py:44
    skip = ctypes.c_int32()
    got = []
    for c in range(case["steps"]):
        st.cnt = c
        L.check(L.lib.mc_ctrl_decide(ctypes.byref(cfg), ctypes.byref(st), ctypes.byref(skip)))
        got.append(str(skip.value))
    assert "".join(got) == case["mask"]


@pytest.mark.parametrize("case", X["qwen_masks"], ids=lambda c: f"{c['table']}-s{c['steps']}-E{c['thresh']}K{c['K']}R{c['R']}")
def test_qwen_image_over_several_images_keeps_accumulators_at_the_wrap(case):
    """Three images and a bit through Qwen-Image's controller, whose wrap statement resets ONLY the counter
    (MagCache4QwenImage/magcache_generate.py:243-244): C ABI (config flag MC_CTRL_WRAP_KEEPS_ACC from the family table), the
    attribute-level shim and the independent Python restatement against the mask and final state the reference's own statements
    produce; without the flag the schedule of the later images must differ for at least one preset (the flag is not a no-op)."""
    from magcache_b200.config import FAMILIES
    from magcache_b200.controller import AttrController
    from oracle.controller_ref import AdapterControllerRef
    tbl = np.array(REMAINING[case["table"]]["values"])
    steps = case["steps"]
    ratios = tbl
    if len(tbl) != 2 * steps:
        def qi(a, T):
            out = np.empty(T)
            dp = ctypes.POINTER(ctypes.c_double)
            a = np.ascontiguousarray(a, dtype=np.float64)
            L.check(L.lib.mc_nearest_interp_linspace(a.ctypes.data_as(dp), len(a), out.ctypes.data_as(dp), T))
            return out
        ratios = np.stack([qi(tbl[0::2], steps), qi(tbl[1::2], steps)], axis=1).reshape(-1)
    kw = FAMILIES["qwen-image"]
    assert kw["flags"] & L.MC_CTRL_WRAP_KEEPS_ACC
    cfg = make_ctrl_config(2 * steps, case["thresh"], case["K"], case["R"], ratios, **kw)
    mask = "".join(map(str, schedule_mask(cfg, case["calls"]).tolist()))
    assert mask == case["mask"]
    ref = AdapterControllerRef("qwen-image", ratios, 2 * steps, case["thresh"], case["K"], case["R"])
    assert "".join(map(str, ref.mask(case["calls"]))) == case["mask"]
    # attribute-level shim (what the patched forward drives)
    import types
    o = types.SimpleNamespace(cnt=0, num_steps=2 * steps, magcache_thresh=case["thresh"], K=case["K"], retention_ratio=case["R"],
                              accumulated_ratio=[1.0, 1.0], accumulated_err=[0.0, 0.0], accumulated_steps=[0, 0], mag_ratios=ratios)
    ctrl = AttrController(kw)
    got = []
    for _ in range(case["calls"]):
        got.append(int(ctrl.decide(o)))
        ctrl.advance(o)
    assert "".join(map(str, got)) == case["mask"]
    assert o.cnt == case["final"]["cnt"] and list(map(float, o.accumulated_err)) == case["final"]["accumulated_err"]
    assert list(map(float, o.accumulated_ratio)) == case["final"]["accumulated_ratio"] and list(map(float, o.accumulated_steps)) == case["final"]["accumulated_steps"]


def test_qwen_wrap_flag_changes_some_schedule():
    from magcache_b200.config import FAMILIES
    kw = dict(FAMILIES["qwen-image"])
    differs = 0
    for case in X["qwen_masks"]:
        tbl = np.array(REMAINING[case["table"]]["values"])
        if len(tbl) != 2 * case["steps"]:
            continue
        with_flag = schedule_mask(make_ctrl_config(2 * case["steps"], case["thresh"], case["K"], case["R"], tbl, **kw), case["calls"])
        without = schedule_mask(make_ctrl_config(2 * case["steps"], case["thresh"], case["K"], case["R"], tbl, **dict(kw, flags=0)), case["calls"])
        differs += int(not np.array_equal(with_flag, without))
    assert differs >= 1
