"""HunyuanVideo engine orchestration on CPU: `magcache_hunyuan_forward` + `HunyuanEngine` through tests/emu_ops.py against the oracle
restatement of MagCache4HunyuanVideo/magcache_sample_video.py:29-160 — fused qkv / linear1 weight splitting, image-first row ranges,
RoPE on the image rows only, the token refiner on the valid text tokens (the padded ones are dropped: they are a separate attention
segment upstream), the one-GEMM modulation table, final layer (shift, scale) and unpatchify order, hit / miss, controller state."""
import copy

import pytest
import torch

import magcache_b200 as mc
from magcache_b200 import mmdit as mmdit_mod
from magcache_b200 import patch as patch_mod
from oracle import hunyuan_ref as hr

import emu_ops


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


@pytest.fixture()
def emulated(monkeypatch):
    monkeypatch.setattr(mmdit_mod, "ops", emu_ops)
    monkeypatch.setattr(patch_mod, "ops", emu_ops)
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True))


def _model(seed=0, guidance=True):
    return hr.HYVideoDiffusionTransformer(hidden_size=256, heads_num=2, mm_double_blocks_depth=2, mm_single_blocks_depth=3, text_states_dim=96,
                                          text_states_dim_2=48, guidance_embed=guidance).init_synthetic(seed)


def _inputs(seed=0, grid=(2, 4, 6), n_txt=16, valid=11):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(1, 16, grid[0], 2 * grid[1], 2 * grid[2], generator=g).bfloat16()
    txt = torch.randn(1, n_txt, 96, generator=g).bfloat16()
    mask = torch.zeros(1, n_txt, dtype=torch.long)
    mask[0, :valid] = 1
    pooled = torch.randn(1, 48, generator=g).bfloat16()
    cos, sin = hr.rope_cos_sin(grid)
    return x, txt, mask, pooled, cos, sin


@pytest.mark.parametrize("valid,guidance", [(11, True), (16, True), (8, False)])
def test_hunyuan_engine_single_forward_matches_oracle(emulated, valid, guidance):
    model = _model(guidance=guidance)
    x, txt, mask, pooled, cos, sin = _inputs(valid=valid)
    t, gd = torch.tensor([731.0]), (torch.tensor([6000.0]) if guidance else None)
    ref_m = copy.deepcopy(model)
    ref_m.__class__ = type("RefHY", (ref_m.__class__,), {})
    hr.install_magcache(type(ref_m), mc.tables()["hunyuan_720p"], 50)
    m64 = copy.deepcopy(model).double()
    m64.__class__ = type("RefHY64", (m64.__class__,), {})
    hr.install_magcache(type(m64), mc.tables()["hunyuan_720p"], 50)
    ours = copy.deepcopy(model)
    ours.__class__ = type("OurHY", (ours.__class__,), {})
    mc.init_magcache_hunyuan(ours, 50)
    assert type(ours).forward is mc.magcache_hunyuan_forward and type(ours).mag_ratios.tolist() == type(ref_m).mag_ratios.tolist()
    with torch.no_grad():
        ref = ref_m(x, t, txt, mask, pooled, cos, sin, gd)["x"]
        with hr.exact():
            exact = m64(x.double(), t.double(), txt.double(), mask, pooled.double(), cos.double(), sin.double(), None if gd is None else gd.double())["x"]
        out = ours(x, t, txt, mask, pooled, cos, sin, gd)["x"]
    assert out.shape == ref.shape == x.shape and out.dtype == torch.bfloat16
    e_ours, e_ref, e_vs = rel_l2(out, exact), rel_l2(ref, exact), rel_l2(out, ref)
    print(f"[hunyuan emulated] ours vs fp64 {e_ours:.3e} | oracle(bf16) vs fp64 {e_ref:.3e} | ours vs oracle {e_vs:.3e}")
    assert e_ours <= 1.5 * e_ref + 1e-3 and e_vs <= 2.0 * e_ref + 1e-3
    assert tuple(ours.residual_cache.shape) == tuple(ref_m.residual_cache.shape)
    assert rel_l2(ours.residual_cache[0], ref_m.residual_cache[0]) <= 2.0 * e_ref + 2e-2


def test_hunyuan_engine_loop_hits_and_misses(emulated):
    model = _model(seed=1)
    x, txt, mask, pooled, cos, sin = _inputs(1)
    steps = 10
    ref_m = copy.deepcopy(model)
    ref_m.__class__ = type("RefHYL", (ref_m.__class__,), {})
    hr.install_magcache(type(ref_m), mc.tables()["hunyuan_720p"], steps, thresh=0.24, K=6, retention_ratio=0.2)
    ours = copy.deepcopy(model)
    ours.__class__ = type("OurHYL", (ours.__class__,), {})
    mc.init_magcache_hunyuan(ours, steps, thresh=0.24, K=6, retention_ratio=0.2)
    want = mc.MagCacheConfig("hunyuan", 0.24, 6, 0.2, steps, table="hunyuan_720p").schedule().tolist()
    skips = []
    with torch.no_grad():
        for i in range(steps + 2):
            t = torch.tensor([1000.0 - 90.0 * (i % steps)])
            xi = x * (1.0 - 0.03 * i)
            ref = ref_m(xi, t, txt, mask, pooled, cos, sin, torch.tensor([6000.0]), return_dict=False)
            out = ours(xi, t, txt, mask, pooled, cos, sin, torch.tensor([6000.0]), return_dict=False)
            skips.append(int(ref_m.last_skip))
            assert rel_l2(out, ref) <= 0.15, (i, rel_l2(out, ref))
            for attr in ("cnt", "accumulated_ratio", "accumulated_err", "accumulated_steps"):
                assert float(getattr(ours, attr)) == float(getattr(ref_m, attr)), (i, attr)
    assert skips[:steps] == want and 0 < sum(want) < steps


def test_hunyuan_forward_argument_checks(emulated):
    ours = _model()
    ours.__class__ = type("OurHYX", (ours.__class__,), {})
    mc.init_magcache_hunyuan(ours, 50)
    x, txt, mask, pooled, cos, sin = _inputs()
    with pytest.raises(ValueError):
        ours(x, torch.tensor([500.0]), txt, mask, pooled, cos, sin, None)  # guidance-distilled model without guidance (:58-61)
    bad = mask.clone()
    bad[0, 0] = 0
    with pytest.raises(NotImplementedError):
        ours(x, torch.tensor([500.0]), txt, bad, pooled, cos, sin, torch.tensor([6000.0]))
    with pytest.raises(KeyError):
        mc.init_magcache_hunyuan(ours, 50, video_height=480)


def test_hunyuan_calibration_twin(emulated, capsys):
    """`magcache_hunyuan_calibration` (magcache_sample_video.py:163-290): always computes (outputs equal the cache-less forward), statistics
    from the second call on against the reference expressions on the two residuals, counter never wrapped (:281)."""
    from oracle.controller_ref import calibration_stats
    model = _model(seed=3)
    x, txt, mask, pooled, cos, sin = _inputs(3)
    ref_m = copy.deepcopy(model)
    ref_m.__class__ = type("RefHYC", (ref_m.__class__,), {})
    hr.install_magcache(type(ref_m), [1.0] * 50, 50, thresh=-1.0)  # negative threshold: the oracle forward never skips
    ours = copy.deepcopy(model)
    ours.__class__ = type("OurHYC", (ours.__class__,), {})
    mc.init_magcache_hunyuan_calibration(ours, 50)
    prev = None
    with torch.no_grad():
        for i in range(4):
            t = torch.tensor([900.0 - 100.0 * i])
            xi = x * (1.0 - 0.1 * i)
            a = ref_m(xi, t, txt, mask, pooled, cos, sin, torch.tensor([6000.0]))["x"]
            b = ours(xi, t, txt, mask, pooled, cos, sin, torch.tensor([6000.0]))["x"]
            assert rel_l2(b, a) <= 2e-2 and not ref_m.last_skip
            cur = ref_m.residual_cache.float()
            if prev is not None:
                want = calibration_stats(cur, prev)[0]
                assert abs(ours.norm_ratio[-1] - want) <= 2e-2 * abs(want), (i, ours.norm_ratio[-1], want)
            prev = cur
    assert ours.cnt == 4 and len(ours.norm_ratio) == len(ours.norm_std) == len(ours.cos_dis) == 3
    assert "time: 1, norm_ratio" in capsys.readouterr().out
