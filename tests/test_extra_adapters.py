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
