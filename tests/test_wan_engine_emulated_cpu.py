"""Wan engine orchestration on CPU: `magcache_forward` / `magcache_vace_forward` + `WanEngine` driven through tests/emu_ops.py (torch
emulation of the kernels' documented arithmetic) against the oracle — weight repacking, workspace views, modulation indices, the i2v
image branch, the VACE control pass and its hint GEMMs, hit / miss branches, residual slots, controller state. Runs in the CPU suite;
the kernels themselves are checked on the GPU by tests/test_wan_forward_gpu.py."""
import copy

import pytest
import torch

import magcache_b200 as mc
from magcache_b200 import patch as patch_mod
from magcache_b200 import wan as wan_mod
from oracle import wan_ref

import emu_ops


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


@pytest.fixture()
def emulated(monkeypatch):
    monkeypatch.setattr(wan_mod, "ops", emu_ops)
    monkeypatch.setattr(patch_mod, "ops", emu_ops)
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True))


def _attach_engine(model):
    object.__setattr__(model, "_mc_engine", mc.WanEngine(mc.WanWeights.from_module(model, torch.device("cpu"))))
    return model


def _pair(model, table, steps, vace=False, **kw):
    ref = copy.deepcopy(model)
    ref.__class__ = type("Ref", (ref.__class__,), {})
    wan_ref.install_magcache(type(ref), table, steps, vace=vace, **kw)
    ours = copy.deepcopy(model)
    ours.__class__ = type("Ours", (ours.__class__,), {})
    mc.init_magcache(ours, steps, mag_ratios=table, **kw)
    return ref, _attach_engine(ours)


KW = dict(thresh=10.0, K=3, retention_ratio=0.25)  # 4 steps = 8 calls: miss, miss, then hits


@pytest.mark.parametrize("kind", ["t2v", "i2v", "vace"])
def test_wan_engine_forward_sequence_matches_oracle(emulated, kind):
    g = torch.Generator().manual_seed(3)
    common = dict(dim=256, ffn_dim=512, num_heads=2, num_layers=4 if kind == "vace" else 2, text_dim=128, text_len=32)
    if kind == "i2v":
        model = wan_ref.WanModel(in_dim=36, model_type="i2v", clip_dim=64, **common).init_synthetic(1)
    elif kind == "vace":
        model = wan_ref.WanModel(model_type="vace", vace_in_dim=24, **common).init_synthetic(1)
    else:
        model = wan_ref.WanModel(**common).init_synthetic(1)
    table = mc.tables()["wan2.1_i2v_480p" if kind == "i2v" else "wan2.1_t2v_1.3b"]
    ref, ours = _pair(model, table, 4, vace=(kind == "vace"), **KW)
    assert type(ours).forward is (mc.magcache_vace_forward if kind == "vace" else mc.magcache_forward)
    lat = torch.randn(16, 2, 8, 12, generator=g)
    ctxs = [torch.randn(19, 128, generator=g), torch.randn(11, 128, generator=g)]
    extra = {}
    if kind == "i2v":
        extra = dict(clip_fea=torch.randn(1, 257, 64, generator=g), y=[torch.randn(20, 2, 8, 12, generator=g)])
    if kind == "vace":
        extra = dict(vace_context=[torch.randn(24, 2, 8, 12, generator=g)], vace_context_scale=0.75)
    n_tok, t = 2 * 4 * 6, torch.tensor([611.0])
    with torch.no_grad():
        for i in range(5):
            a = ref([lat], t=t, context=[ctxs[i % 2]], seq_len=n_tok, **extra)[0]
            b = ours([lat], t=t, context=[ctxs[i % 2]], seq_len=n_tok, **extra)[0]
            assert bool(ref.last_skip) == (i >= 2)
            assert b.shape == a.shape and rel_l2(b, a) <= 1.5e-2, (kind, i, rel_l2(b, a))
            assert ours.cnt == ref.cnt and ours.accumulated_err == ref.accumulated_err and ours.accumulated_steps == ref.accumulated_steps
            slot = (ref.cnt - 1) % 2
            assert rel_l2(ours.residual_cache[slot][0], ref.residual_cache[slot][0]) <= 3e-2


def test_wan_engine_calibration_and_eval_variant(emulated, tmp_path):
    model = wan_ref.WanModel(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=128, text_len=32).init_synthetic(2)
    ref = copy.deepcopy(model)
    ref.__class__ = type("RefC", (ref.__class__,), {})
    wan_ref.install_magcache(type(ref), None, 3, calibration=True)
    ours = copy.deepcopy(model)
    ours.__class__ = type("OursC", (ours.__class__,), {})
    mc.init_magcache_calibration(ours, 3)
    type(ours).calibration_dir = str(tmp_path)
    _attach_engine(ours)
    g = torch.Generator().manual_seed(5)
    lat, ctx = torch.randn(16, 2, 8, 8, generator=g), torch.randn(9, 128, generator=g)
    with torch.no_grad():
        for i in range(6):
            x = lat * (1.0 - 0.1 * i)
            ref([x], t=torch.tensor([900.0 - 50 * i]), context=[ctx], seq_len=32)
            ours([x], t=torch.tensor([900.0 - 50 * i]), context=[ctx], seq_len=32)
    assert len(ours.norm_ratio) == len(ref.norm_ratio) == 4
    for a, b in zip(ours.norm_ratio, ref.norm_ratio):
        assert abs(a - b) <= 2e-2 * abs(b)
    assert (tmp_path / "wan2_1_mag_ratio.json").exists()


def test_wan_engine_calibration_with_padded_seq_len(emulated, tmp_path):
    """seq_len > token count (MagCache4Wan2.1/magcache_generate.py:243-246 pads the sequence with zero rows): the reference's three
    statistics (:167-169) average over the padded rows too. The engine computes ONE representative pad row (they are identical:
    zero input, no RoPE, same keys) as an extra query and weights it by the pad count — same means and std as the oracle, which
    really carries all the padded rows."""
    model = wan_ref.WanModel(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=128, text_len=32).init_synthetic(4)
    ref = copy.deepcopy(model)
    ref.__class__ = type("RefCP", (ref.__class__,), {})
    wan_ref.install_magcache(type(ref), None, 3, calibration=True)
    ours = copy.deepcopy(model)
    ours.__class__ = type("OursCP", (ours.__class__,), {})
    mc.init_magcache_calibration(ours, 3)
    type(ours).calibration_dir = str(tmp_path)
    _attach_engine(ours)
    g = torch.Generator().manual_seed(6)
    lat, ctx = torch.randn(16, 2, 8, 8, generator=g), torch.randn(9, 128, generator=g)
    n_tok, seq_len = 32, 41  # 9 padded rows = 22 % of the sequence: the statistics differ visibly from the unpadded ones
    outs = []
    with torch.no_grad():
        for i in range(6):
            x = lat * (1.0 - 0.1 * i)
            a = ref([x], t=torch.tensor([900.0 - 50 * i]), context=[ctx], seq_len=seq_len)[0]
            b = ours([x], t=torch.tensor([900.0 - 50 * i]), context=[ctx], seq_len=seq_len)[0]
            outs.append(rel_l2(b, a))
    assert max(outs) <= 1.5e-2, outs
    assert len(ours.norm_ratio) == len(ref.norm_ratio) == 4
    for name in ("norm_ratio", "norm_std", "cos_dis"):
        for a, b in zip(getattr(ours, name), getattr(ref, name)):
            assert abs(a - b) <= 3e-2 * abs(b) + 2e-4, (name, a, b)
    assert ours._mc_engine.pad_row == 1 and ours.residual_cache[0].shape[1] == n_tok + 1


def test_wan_engine_calibration_with_per_token_timesteps_and_padding(emulated, tmp_path):
    """Calibration of a TI2V-5B-shaped model (MagCache4Wan2.2/magcache_generate.py:80-194 twin: 48 channels, `t` [1, seq_len] with the
    first-frame tokens at t = 0, seq_len > token count): the representative pad row takes the padded positions' timestep, the row
    ranges stop at the tokens in the head, and the statistics equal the oracle's (which carries every padded row)."""
    model = wan_ref.WanModel(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=128, text_len=32, in_dim=48, out_dim=48).init_synthetic(5)
    ref = copy.deepcopy(model)
    ref.__class__ = type("RefCT", (ref.__class__,), {})
    wan_ref.install_magcache(type(ref), None, 3, calibration=True)
    ours = copy.deepcopy(model)
    ours.__class__ = type("OursCT", (ours.__class__,), {})
    mc.init_magcache_calibration(ours, 3)
    type(ours).calibration_dir = str(tmp_path)
    _attach_engine(ours)
    g = torch.Generator().manual_seed(8)
    lat, ctx = torch.randn(48, 2, 8, 8, generator=g), torch.randn(9, 128, generator=g)
    n_tok, seq_len = 32, 39
    outs = []
    with torch.no_grad():
        for i in range(6):
            t = torch.full((1, seq_len), 900.0 - 50 * i)
            t[0, :16] = 0.0
            x = lat * (1.0 - 0.1 * i)
            a = ref([x], t=t, context=[ctx], seq_len=seq_len)[0]
            b = ours([x], t=t, context=[ctx], seq_len=seq_len)[0]
            assert b.shape == (48, 2, 8, 8)
            outs.append(rel_l2(b, a))
    assert max(outs) <= 1.5e-2, outs
    for name in ("norm_ratio", "norm_std", "cos_dis"):
        for a, b in zip(getattr(ours, name), getattr(ref, name)):
            assert abs(a - b) <= 3e-2 * abs(b) + 2e-4, (name, a, b)
    eng = ours._mc_engine
    assert eng.pad_row == 1 and eng.runs == [(0, 16, 0), (16, 33, 1)] and ours.residual_cache[0].shape[1] == n_tok + 1
    with pytest.raises(NotImplementedError):  # padded positions with different timesteps: no single representative row
        t = torch.full((1, seq_len), 500.0)
        t[0, -1] = 0.0
        ours([lat], t=t, context=[ctx], seq_len=seq_len)
