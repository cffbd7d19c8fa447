"""Wan2.2 A14B (MagCache4Wan2.2/magcache_generate.py:209-362) on the Wan engine, CPU with emulated kernels: TWO expert instances of one
class — high-noise model for the first `split_step` calls, low-noise model afterwards — sharing the tensor counter, the accumulators and
the residual-cache list exactly like the reference, each with its own weights / engine. Same hit / miss sequence as the oracle (and as the
golden Wan2.2 windows), outputs to bf16-pipeline noise, residual hand-over between the experts."""
import copy

import pytest
import torch

import magcache_b200 as mc
from magcache_b200 import patch as patch_mod
from magcache_b200 import wan as wan_mod
from oracle import wan_ref

import emu_ops


@pytest.fixture()
def emulated(monkeypatch):
    monkeypatch.setattr(wan_mod, "ops", emu_ops)
    monkeypatch.setattr(patch_mod, "ops", emu_ops)
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True))


@pytest.mark.parametrize("mode", ["t2v", "i2v"])
def test_two_experts_share_controller_and_cache(emulated, mode):
    kw = dict(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=128, text_len=32)
    if mode == "i2v":
        kw.update(in_dim=36, model_type="i2v")  # Wan2.2 I2V: y under the latents, no CLIP branch
    protos = []
    for seed in (1, 2):
        m = wan_ref.WanModel(**{k: v for k, v in kw.items() if k != "model_type"}).init_synthetic(seed)
        m.model_type = kw.get("model_type", "t2v")
        protos.append(m)
    steps, high = 12, 5  # int(split_step * R) = 2: both cache slots are filled before the first eligible call
    ratios = mc.tables()["wan2.2_i2v_a14b" if mode == "i2v" else "wan2.2_t2v_a14b"][2:].tolist()
    RefCls = type("RefW22", (wan_ref.WanModel,), {})
    OurCls = type("OurW22", (wan_ref.WanModel,), {})
    refs, ours = [], []
    for p in protos:
        r, o = copy.deepcopy(p), copy.deepcopy(p)
        r.__class__, o.__class__ = RefCls, OurCls
        object.__setattr__(o, "_mc_engine", mc.WanEngine(mc.WanWeights.from_module(o, torch.device("cpu"))))
        refs.append(r)
        ours.append(o)
    wan_ref.install_magcache_wan22(RefCls, ratios, steps, thresh=0.12, K=2, retention_ratio=0.2, split_steps=high, mode=mode)
    mc.init_magcache_wan22(ours[0], ratios, steps, thresh=0.12, K=2, retention_ratio=0.2, split_steps=high, mode=mode)
    assert OurCls.mag_ratios.tolist() == RefCls.mag_ratios.tolist() and OurCls.split_step == 2 * high
    g = torch.Generator().manual_seed(0)
    lat = torch.randn(16, 2, 8, 8, generator=g)
    ys = [torch.randn(20, 2, 8, 8, generator=g)] if mode == "i2v" else None
    ctxs = [torch.randn(9, 128, generator=g), torch.randn(7, 128, generator=g)]
    kinds = []
    with torch.no_grad():
        for call in range(2 * steps):
            e = 0 if call < 2 * high else 1  # which expert the pipeline calls (boundary timestep)
            t = torch.tensor([950.0 - 40.0 * (call // 2)])
            a = refs[e]([lat], t=t, context=[ctxs[call % 2]], seq_len=32, y=ys)[0]
            b = ours[e]([lat], t=t, context=[ctxs[call % 2]], seq_len=32, y=ys)[0]
            kinds.append(int(refs[e].last_skip))
            rel = float((a - b).norm() / a.norm())
            assert rel < 2e-2, (mode, call, rel)
            if call < 2 * steps - 1:
                assert int(OurCls.cnt) == int(RefCls.cnt) == call + 1 and torch.is_tensor(OurCls.cnt)  # ONE counter for both experts
            assert OurCls.accumulated_err == RefCls.accumulated_err and OurCls.accumulated_steps == RefCls.accumulated_steps
    fam = "wan2.2-i2v" if mode == "i2v" else "wan2.2-t2v"
    want = mc.MagCacheConfig(fam, 0.12, 2, 0.2, steps, mag_ratios=OurCls.mag_ratios, high_noise_steps=high).schedule().tolist()
    assert kinds == want and 0 < sum(want) < 2 * steps
    assert any(kinds[2 * high:]), "the low-noise expert must reach hits (its residual comes from its own misses after the window)"
    assert int(OurCls.cnt) == 0


def test_per_token_timesteps_are_refused(emulated):
    m = wan_ref.WanModel(dim=256, ffn_dim=512, num_heads=2, num_layers=1, text_dim=128, text_len=32).init_synthetic(0)
    m.__class__ = type("OurTI2V", (wan_ref.WanModel,), {})
    object.__setattr__(m, "_mc_engine", mc.WanEngine(mc.WanWeights.from_module(m, torch.device("cpu"))))
    mc.init_magcache_wan22(m, mc.tables()["wan2.2_ti2v_5b_a"][2:].tolist(), 50)
    t = torch.full((1, 32), 500.0)
    t[0, :8] = 0.0
    with pytest.raises(NotImplementedError):
        m([torch.randn(16, 2, 8, 8)], t=t, context=[torch.randn(5, 128)], seq_len=32)
    out = m([torch.randn(16, 2, 8, 8)], t=torch.full((1, 32), 500.0), context=[torch.randn(5, 128)], seq_len=32)[0]  # uniform [B, L] is fine
    assert out.shape == (16, 2, 8, 8)
