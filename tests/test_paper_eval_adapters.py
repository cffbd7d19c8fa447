"""SURVEY Appendix A — the remaining controller variants, bit-exact against fixtures produced by executing the reference's own
statements (tests/golden/make_golden.py::paper_eval_cases):

* FramePack / FramePack-F1 (MagCache4FramePack/magcache_demo_gradio.py:252-270, :298-300): scalar state re-initialised whenever a
  call sees cnt == 0, `cnt >= 1` guard, and the per-step veto `|1 - mag_ratio| <= 0.06`;
* the paper-evaluation Wan2.1 forward (eval/magcache/experiments/Wan2.1_EVAL/wan_magcache.py:770-786, :807-815): `<=`, table
  indexed `ratio[t - 10]` (90 entries for 100 calls);
* the paper-evaluation Open-Sora forward (eval/magcache/experiments/opensora.py:297-308, :348-354): explicit `skip_time`,
  `ratio[t - 1]`, error accumulated WITHOUT abs.
"""
import ctypes
import json
import os

import numpy as np
import pytest

from magcache_b200 import _lib as L
from magcache_b200.controller import make_ctrl_config, schedule_mask

with open(os.path.join(os.path.dirname(__file__), "golden", "paper_eval_adapters.json")) as f:
    P = json.load(f)


def _stepwise(cfg, calls):
    st = L.CtrlState()
    st.accumulated_ratio[0] = st.accumulated_ratio[1] = 1.0
    skip = ctypes.c_int32()
    got = []
    for _ in range(calls):
        L.check(L.lib.mc_ctrl_decide(ctypes.byref(cfg), ctypes.byref(st), ctypes.byref(skip)))
        got.append(str(skip.value))
        L.check(L.lib.mc_ctrl_advance(ctypes.byref(cfg), ctypes.byref(st)))
    return "".join(got), st


def framepack_cfg(case):
    return make_ctrl_config(case["steps"], case["thresh"], case["K"], case["R"], case["ratios"], branches=1, cmp=L.MC_CMP_LE,
                            retention_mode=L.MC_RETAIN_FLOOR, min_cnt=1, flags=L.MC_CTRL_RESET_AT_ZERO, ratio_veto=0.06)


@pytest.mark.parametrize("case", P["framepack_masks"], ids=lambda c: f"{c['table']}-s{c['steps']}-E{c['thresh']}K{c['K']}R{c['R']}")
def test_framepack(case):
    from magcache_b200.config import nearest_interp
    tbl = np.array(P["tables"][case["table"]]["values"])
    ratios = tbl if len(tbl) == case["steps"] else nearest_interp(tbl, case["steps"])  # initialize_magcache :72-74
    assert ratios.tolist() == case["ratios"]
    cfg = framepack_cfg(case)
    assert "".join(map(str, schedule_mask(cfg, case["calls"]).tolist())) == case["mask"]
    mask, st = _stepwise(cfg, case["calls"])
    assert mask == case["mask"]
    assert st.cnt == case["final"]["cnt"] != 0
    assert st.accumulated_err[0] == case["final"]["accumulated_err"]
    assert st.accumulated_ratio[0] == case["final"]["accumulated_ratio"]
    assert st.accumulated_steps[0] == case["final"]["accumulated_steps"]


def test_framepack_reinitialises_when_cnt_is_reset_from_outside():
    """The gradio demo re-creates the state with `initialize_magcache` between sections (cnt = 0 mid-video): the accumulators of the
    interrupted section must not leak into the next one (magcache_demo_gradio.py:253-256)."""
    case = [c for c in P["framepack_masks"] if c["steps"] == 25 and c["thresh"] == 0.1 and c["K"] == 3 and c["table"] == "framepack"][0]
    cfg = framepack_cfg(case)
    st = L.CtrlState()
    st.accumulated_ratio[0] = st.accumulated_ratio[1] = 1.0
    skip = ctypes.c_int32()
    for _ in range(9):  # stop in the middle of a run of skips
        L.check(L.lib.mc_ctrl_decide(ctypes.byref(cfg), ctypes.byref(st), ctypes.byref(skip)))
        L.check(L.lib.mc_ctrl_advance(ctypes.byref(cfg), ctypes.byref(st)))
    assert st.accumulated_steps[0] > 0 or st.accumulated_err[0] > 0
    st.cnt = 0
    got = []
    for _ in range(25):
        L.check(L.lib.mc_ctrl_decide(ctypes.byref(cfg), ctypes.byref(st), ctypes.byref(skip)))
        got.append(str(skip.value))
        L.check(L.lib.mc_ctrl_advance(ctypes.byref(cfg), ctypes.byref(st)))
    assert "".join(got) == case["mask"][:25]


@pytest.mark.parametrize("case", P["eval_wan_masks"], ids=lambda c: f"E{c['thresh']}K{c['K']}")
def test_eval_wan_table_offset(case):
    tbl = np.array(P["tables"]["wan2.1_eval"]["values"])
    assert len(tbl) == 90
    cfg = make_ctrl_config(2 * case["steps"], case["thresh"], case["K"], 0.2, tbl, branches=2, cmp=L.MC_CMP_LE, retention_mode=L.MC_RETAIN_FLOOR,
                           table_offset=10)
    L.check(L.lib.mc_ctrl_validate(ctypes.byref(cfg)))
    mask, st = _stepwise(cfg, case["calls"])
    assert mask == case["mask"]
    assert sum(map(int, mask[:100])) == case["skipped_first_video"]
    assert st.cnt == case["final"]["cnt"]
    for i in range(2):
        assert st.accumulated_err[i] == case["final"]["accumulated_err"][i]
        assert st.accumulated_ratio[i] == case["final"]["accumulated_ratio"][i]
        assert st.accumulated_steps[i] == case["final"]["accumulated_steps"][i]


def test_eval_wan_paper_presets_skip_counts():
    """"slow" = 0.12/K2 and "fast" = 0.12/K4 of eval/magcache/experiments/Wan2.1_EVAL/wan_eval.sh:30-31, 66-67."""
    by = {(c["thresh"], c["K"]): c for c in P["eval_wan_masks"]}
    assert by[(0.12, 2)]["skipped_first_video"] == 52 and by[(0.12, 4)]["skipped_first_video"] == 62
    assert by[(0.12, 2)]["mask"][:20] == "0" * 20  # int(100 * 0.2) retained calls


@pytest.mark.parametrize("case", P["opensora_masks"], ids=lambda c: f"E{c['thresh']}K{c['K']}T{c['skip_time']}")
def test_opensora_signed_error(case):
    tbl = np.array(P["tables"]["opensora_eval"]["values"])
    assert len(tbl) == 29
    cfg = make_ctrl_config(30, case["thresh"], case["K"], 0.0, tbl, branches=1, cmp=L.MC_CMP_LE, retention_mode=L.MC_RETAIN_EXPLICIT,
                           split_step=case["skip_time"], table_offset=1, flags=L.MC_CTRL_SIGNED_ERR)
    mask, st = _stepwise(cfg, case["calls"])
    assert mask == case["mask"]
    assert st.cnt == case["final"]["cnt"]
    assert st.accumulated_err[0] == case["final"]["accumulated_err"]
    assert st.accumulated_ratio[0] == case["final"]["accumulated_ratio"]
    assert st.accumulated_steps[0] == case["final"]["accumulated_steps"]


def test_signed_error_differs_from_abs_when_ratios_exceed_one():
    """With ratios > 1 the signed error goes negative and never reaches the threshold; the abs form does."""
    ratios = np.full(20, 1.05)
    kw = dict(branches=1, cmp=L.MC_CMP_LE, retention_mode=L.MC_RETAIN_EXPLICIT, split_step=2)
    signed = schedule_mask(make_ctrl_config(20, 0.12, 100, 0.0, ratios, flags=L.MC_CTRL_SIGNED_ERR, **kw), 20)
    absd = schedule_mask(make_ctrl_config(20, 0.12, 100, 0.0, ratios, **kw), 20)
    assert signed[2:].all() and not absd[2:].all()


def test_table_offset_before_retention_window_is_rejected():
    """num_steps < 50 in the eval forward makes `ratio[t-10]` a negative (wrap-around) index upstream; here it is an error."""
    tbl = np.full(30, 0.99)
    cfg = make_ctrl_config(40, 0.12, 2, 0.2, tbl, branches=2, cmp=L.MC_CMP_LE, retention_mode=L.MC_RETAIN_FLOOR, table_offset=10)
    assert L.lib.mc_ctrl_validate(ctypes.byref(cfg)) == L.MC_ERR_STATE
    assert b"table_offset" in L.lib.mc_last_error()
    with pytest.raises(L.MagCacheError):
        schedule_mask(cfg, 40)


def test_unknown_flag_bits_and_short_tables_are_rejected():
    ratios = np.full(10, 0.99)
    cfg = make_ctrl_config(10, 0.1, 2, 0.2, ratios, branches=1, cmp=L.MC_CMP_LE, retention_mode=L.MC_RETAIN_FLOOR)
    cfg.flags = 64
    assert L.lib.mc_ctrl_validate(ctypes.byref(cfg)) == L.MC_ERR_INVALID
    with pytest.raises(IndexError):
        make_ctrl_config(10, 0.1, 2, 0.2, np.full(8, 0.99), branches=1, cmp=L.MC_CMP_LE, retention_mode=L.MC_RETAIN_FLOOR, table_offset=1)


def test_presets_reproduce_the_reference_schedules():
    """`MagCacheConfig.schedule()` of the shipped presets against the masks produced by the reference's statements."""
    import magcache_b200 as mc
    with open(os.path.join(os.path.dirname(__file__), "golden", "extra_adapters.json")) as f:
        X = json.load(f)
    with open(os.path.join(os.path.dirname(__file__), "golden", "masks.json")) as f:
        M = json.load(f)

    def sched(name):
        return "".join(map(str, mc.PRESETS[name].schedule().tolist()))

    ew = {(c["thresh"], c["K"]): c["mask"][:100] for c in P["eval_wan_masks"]}
    assert sched("wan2.1-eval-slow-E012K2") == ew[(0.12, 2)] and sched("wan2.1-eval-fast-E012K4") == ew[(0.12, 4)]
    osr = {(c["thresh"], c["K"], c["skip_time"]): c["mask"][:30] for c in P["opensora_masks"]}
    assert sched("opensora-slow-E012K3") == osr[(0.12, 3, 6)] and sched("opensora-fast-E024K5") == osr[(0.24, 5, 6)]
    fp = {(c["table"], c["steps"], c["thresh"], c["K"], c["R"]): c["mask"][:c["steps"]] for c in P["framepack_masks"]}
    assert sched("framepack-E010K3R02") == fp[("framepack", 25, 0.1, 3, 0.2)]
    assert sched("framepack-f1-E010K3R02") == fp[("framepack_f1", 25, 0.1, 3, 0.2)]
    om = {(c["table"], c["steps"], c["thresh"], c["K"], c["R"]): c["mask"] for c in X["omnigen2_masks"]}
    assert sched("omnigen2-t2i-cond-E006K3R02") == om[("omnigen2_t2i_cond", 50, 0.06, 3, 0.2)]
    w22 = {(c["table"], c["mode"], c["steps"], c["high_noise_steps"], c["thresh"], c["K"], c["R"]): c["mask"][:2 * c["steps"]] for c in X["wan22_masks"]}
    assert sched("wan2.2-ti2v-5b-E006K2R02") == w22[("wan2.2_ti2v_5b_a", "ti2v", 50, None, 0.06, 2, 0.2)]
    assert sched("wan2.2-t2v-a14b-E006K2R04") == mc.MagCacheConfig("wan2.2-t2v", 0.06, 2, 0.4, 40, table="wan2.2_t2v_a14b", high_noise_steps=13).schedule().tolist().__str__().replace(", ", "")[1:-1]
    ms = {(c["family"], c["table"], c["steps"], c["thresh"], c["K"], c["R"]): c for c in M}
    for name, key, n in [("wan2.1-1.3b-E012K4R02", ("wan2.1", "wan2.1_t2v_1.3b", 50, 0.12, 4, 0.2), 100),
                         ("wan2.1-14b-E024K6R02", ("wan2.1", "wan2.1_t2v_14b", 50, 0.24, 6, 0.2), 100),
                         ("wan2.1-vace-1.3b-E002K3R02", ("wan2.1", "wan2.1_vace_1.3b", 50, 0.02, 3, 0.2), 100),
                         ("flux-E024K5R01", ("flux", "flux_dev", 28, 0.24, 5, 0.1), 28),
                         ("hunyuan-720p-E024K6R02", ("hunyuan", "hunyuan_720p", 50, 0.24, 6, 0.2), 50)]:
        assert sched(name) == ms[key]["mask"][:n], name


def test_wan22_expert_presets_match_golden_when_case_exists():
    import magcache_b200 as mc
    with open(os.path.join(os.path.dirname(__file__), "golden", "extra_adapters.json")) as f:
        X = json.load(f)
    hit = 0
    for c in X["wan22_masks"]:
        if c["mode"] not in ("t2v", "i2v") or c["high_noise_steps"] is None:
            continue
        cfg = mc.MagCacheConfig(f"wan2.2-{c['mode']}", c["thresh"], c["K"], c["R"], c["steps"], table=c["table"], high_noise_steps=c["high_noise_steps"])
        assert "".join(map(str, cfg.schedule(c["calls"]).tolist())) == c["mask"]
        hit += 1
    assert hit > 50


def test_shipped_tables_are_the_extracted_reference_literals():
    import magcache_b200 as mc
    with open(os.path.join(os.path.dirname(__file__), "golden", "tables_all.json")) as f:
        T = json.load(f)
    assert set(T) == set(mc.tables())
    for k, v in T.items():
        assert mc.tables()[k].tolist() == v["values"], k
    for name, cfg in mc.PRESETS.items():
        assert cfg.table in T, name
