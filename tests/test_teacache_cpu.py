"""TeaCache comparator host logic — `mc_tea_decide` / `mc_tea_advance` / `mc_tea_needs_distance` bit-exact
against fixtures produced by executing the reference's own controller statements
(eval/magcache/experiments/Wan2.1_EVAL/wan_teacache.py:533-564, :587-589; tests/golden/make_golden.py::teacache_cases) on the
four coefficient sets of its t2v installation block (:913-926)."""
import ctypes
import json
import os

import pytest

from magcache_b200 import _lib as L
from magcache_b200.patch import TEACACHE_COEFFICIENTS

with open(os.path.join(os.path.dirname(__file__), "golden", "teacache.json")) as f:
    T = json.load(f)


def _cfg(case):
    coef = T["coefficients"][case["key"]]
    cfg = L.TeaConfig()
    cfg.num_steps, cfg.ret_steps, cfg.cutoff_steps, cfg.n_coef, cfg.thresh = 2 * case["steps"], case["ret_steps"], case["cutoff_steps"], len(coef), case["thresh"]
    for i, c in enumerate(coef):
        cfg.coef[i] = c
    return cfg


def test_shipped_coefficients_are_the_reference_literals():
    assert TEACACHE_COEFFICIENTS[(True, "1.3B")] == T["coefficients"]["ret_1.3B"]
    assert TEACACHE_COEFFICIENTS[(True, "14B")] == T["coefficients"]["ret_14B"]
    assert TEACACHE_COEFFICIENTS[(False, "1.3B")] == T["coefficients"]["noret_1.3B"]
    assert TEACACHE_COEFFICIENTS[(False, "14B")] == T["coefficients"]["noret_14B"]


@pytest.mark.parametrize("case", T["cases"], ids=lambda c: f"{c['key']}-s{c['steps']}-a{c['amp']}-E{c['thresh']}")
def test_decisions_and_accumulators_bit_exact(case):
    cfg = _cfg(case)
    st = L.TeaState()
    calc, needs = ctypes.c_int32(), ctypes.c_int32()
    got = []
    for d in case["rel_l1"]:
        L.check(L.lib.mc_tea_needs_distance(ctypes.byref(cfg), ctypes.byref(st), ctypes.byref(needs)))
        assert bool(needs.value) == (d is not None)
        L.check(L.lib.mc_tea_decide(ctypes.byref(cfg), ctypes.byref(st), 0.0 if d is None else d, ctypes.byref(calc)))
        got.append(str(calc.value))
        L.check(L.lib.mc_tea_advance(ctypes.byref(cfg), ctypes.byref(st)))
    assert "".join(got) == case["calc"]
    assert st.cnt == case["final"]["cnt"]
    assert st.accumulated[0] == case["final"]["even"] and st.accumulated[1] == case["final"]["odd"]


def test_bad_arguments():
    cfg = _cfg(T["cases"][0])
    st = L.TeaState()
    calc = ctypes.c_int32()
    st.cnt = cfg.num_steps
    assert L.lib.mc_tea_decide(ctypes.byref(cfg), ctypes.byref(st), 0.1, ctypes.byref(calc)) == L.MC_ERR_STATE
    st.cnt = 0
    cfg.n_coef = 9
    assert L.lib.mc_tea_decide(ctypes.byref(cfg), ctypes.byref(st), 0.1, ctypes.byref(calc)) == L.MC_ERR_INVALID
