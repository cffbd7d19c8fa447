"""CPU tests of the Python configuration layer: table selection (the reference's ckpt_dir substring rules), presets,
interpolation through the C ABI vs the oracle, controller attribute plumbing (instance vs class attributes)."""
import numpy as np
import pytest

import magcache_b200 as mc
from magcache_b200.config import interp_cfg, nearest_interp, table_for_ckpt_dir, tables
from magcache_b200.controller import AttrController
from oracle.controller_ref import ControllerRef
from oracle.controller_ref import interp_cfg as ref_interp_cfg
from oracle.controller_ref import nearest_interp as ref_nearest_interp


def test_table_selection_follows_ckpt_dir_substrings():
    """MagCache4Wan2.1/magcache_generate.py:909-912 (t2v), :1001-1004 (i2v), :1141-1144 (VACE)."""
    t = tables()
    assert table_for_ckpt_dir("/w/Wan2.1-T2V-1.3B") is t["wan2.1_t2v_1.3b"]
    assert table_for_ckpt_dir("/w/Wan2.1-T2V-14B") is t["wan2.1_t2v_14b"]
    assert table_for_ckpt_dir("/w/Wan2.1-I2V-14B-480P") is t["wan2.1_i2v_480p"]
    assert table_for_ckpt_dir("/w/Wan2.1-I2V-14B-720P") is t["wan2.1_i2v_720p"]
    assert table_for_ckpt_dir("/w/Wan2.1-VACE-1.3B") is t["wan2.1_vace_1.3b"]
    with pytest.raises(KeyError):
        table_for_ckpt_dir("/w/unknown")
    assert len(t["wan2.1_t2v_1.3b"]) == 100 and t["wan2.1_t2v_1.3b"][0] == 1.0 and len(t["flux_dev"]) == 28


@pytest.mark.parametrize("steps", [50, 40, 30, 25, 20, 10])
def test_interp_cfg_matches_oracle(steps):
    src = tables()["wan2.1_t2v_14b"]
    assert np.array_equal(interp_cfg(src, steps), ref_interp_cfg(src, steps))
    assert np.array_equal(nearest_interp(src[0::2], steps), ref_nearest_interp(src[0::2], steps))


def test_presets_resolve_and_validate():
    for name, cfg in mc.PRESETS.items():
        r = cfg.resolved_ratios()
        off = mc.config.FAMILIES[cfg.family].get("table_offset", 0)
        assert len(r) == cfg.num_steps - off, name
        assert cfg.branches == (2 if cfg.family.startswith(("wan", "qwen")) else 1)
        m = cfg.schedule()
        assert len(m) == cfg.num_steps and 0 < int(m.sum()) < cfg.num_steps and not m[:cfg.branches].any(), name


def test_attr_controller_instance_vs_class_attributes():
    """The reference's `self.cnt += 1` creates an INSTANCE attribute over the class-level one; the list accumulators are
    mutated in place on the class until the end-of-video rebind. The shim reproduces both, and reset_magcache clears both."""
    cfg = mc.PRESETS["wan2.1-1.3b-E012K4R02"]

    class M:
        pass

    m = M()
    mc.init_magcache(m, cfg.sample_steps, thresh=cfg.thresh, K=cfg.K, retention_ratio=cfg.retention_ratio, table="wan2.1_t2v_1.3b")
    assert "cnt" not in m.__dict__ and M.cnt == 0
    ctl = AttrController(dict(branches=2, cmp=0, retention_mode=0, veto_index=-1, veto_base=0))
    ref = ControllerRef("wan2.1", cfg.resolved_ratios(), cfg.num_steps, cfg.thresh, cfg.K, cfg.retention_ratio)
    class_lists = (M.accumulated_ratio, M.accumulated_err, M.accumulated_steps)
    for i in range(cfg.num_steps + 7):
        skip = ctl.decide(m)
        ctl.advance(m)
        assert skip == ref.step(), i
        assert m.cnt == ref.cnt
        if i < cfg.num_steps - 1:
            assert m.accumulated_ratio is class_lists[0]  # mutated in place, as `self.accumulated_ratio[i] = ...` does
            assert list(m.accumulated_ratio) == ref.ratio and list(m.accumulated_err) == ref.err
    assert "cnt" in m.__dict__ and M.cnt == 0            # instance attribute shadows the class one
    assert m.accumulated_ratio is not class_lists[0]     # rebound at the end of the first video (:308-311)
    mc.reset_magcache(m)
    assert "cnt" not in m.__dict__ and m.cnt == 0 and m.accumulated_steps == [0, 0]


def test_handle_classes_do_not_share_state():
    """class-level state makes two pipelines in one process corrupt each other; every WanModelHandle is its own class."""
    class FakeW:
        class dims:
            dim, num_heads = 256, 2
        device = "cpu"
    a, b = mc.WanModelHandle.__new__(mc.WanModelHandle, FakeW), mc.WanModelHandle.__new__(mc.WanModelHandle, FakeW)
    assert type(a) is not type(b) and issubclass(type(a), mc.WanModelHandle)
    type(a).cnt = 5
    assert not hasattr(type(b), "cnt") or type(b).cnt != 5


def test_calibration_dump_round_trips_into_a_table(tmp_path):
    """magcache_generate.py:36-38 / :191-193 file format and the `[1.0]*2 + [...]` table convention (:910-912)."""
    import json

    from magcache_b200.config import save_json, table_from_calibration
    ratios = [1.0124, 1.02213, 1.00166, 1.0041, 0.99791, 1.00061]
    save_json(tmp_path / "wan2_1_mag_ratio", ratios)
    with open(tmp_path / "wan2_1_mag_ratio.json") as f:
        assert json.load(f) == ratios
    t = table_from_calibration(tmp_path / "wan2_1_mag_ratio.json")
    assert t.tolist() == [1.0, 1.0] + ratios
    assert table_from_calibration(ratios, branches=1).tolist() == [1.0] + ratios
    # the shipped 1.3B table is exactly such a dump behind [1.0]*2
    full = mc.tables()["wan2.1_t2v_1.3b"]
    assert table_from_calibration(full[2:].tolist()).tolist() == full.tolist()
    import pytest
    with pytest.raises(ValueError):
        table_from_calibration([])
    with pytest.raises(ValueError):
        table_from_calibration([1.0, float("nan")])


def test_attr_controller_framepack_state_created_on_first_call():
    """FramePack's `initialize_magcache` (magcache_demo_gradio.py:63-74) sets cnt / num_steps / thresh / K / retention / mag_ratios only;
    the accumulators appear when the forward first sees cnt == 0 (:253-256). The attribute shim must start from that state."""
    from magcache_b200.controller import AttrController
    cfg = mc.PRESETS["framepack-E010K3R02"]
    o = type("FP", (), {})()
    o.cnt, o.num_steps, o.magcache_thresh, o.K, o.retention_ratio, o.mag_ratios = 0, cfg.num_steps, cfg.thresh, cfg.K, cfg.retention_ratio, cfg.resolved_ratios()
    ctrl = AttrController(mc.FAMILIES["framepack"])
    got = []
    for _ in range(cfg.num_steps):
        got.append(int(ctrl.decide(o)))
        ctrl.advance(o)
    assert got == cfg.schedule().tolist() and o.cnt == 0
    assert hasattr(o, "accumulated_ratio") and hasattr(o, "accumulated_err") and hasattr(o, "accumulated_steps")
