"""FLUX engine orchestration on CPU: `magcache_flux_forward` + `FluxEngine` driven through tests/emu_ops.py (torch emulation of the
kernels' documented arithmetic) against the oracle restatement of MagCache4FLUX/magcache_flux.py:234-440 — weight packing from the
diffusers attribute names, the one-GEMM AdaLayerNorm table and its offsets, text-first row ranges, q|k / V^T / cat buffer views, RoPE
table, gates, hit / miss branches, residual cache, controller state. The kernels themselves need a GPU (tests/test_flux_forward_gpu.py)."""
import copy

import pytest
import torch

import magcache_b200 as mc
from magcache_b200 import mmdit as flux_mod
from magcache_b200 import patch as patch_mod
from oracle import flux_ref as fr

import emu_ops


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


@pytest.fixture()
def emulated(monkeypatch):
    monkeypatch.setattr(flux_mod, "ops", emu_ops)
    monkeypatch.setattr(patch_mod, "ops", emu_ops)
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True))  # the forward insists on CUDA tensors


def _model(guidance=True, seed=0):
    return fr.FluxTransformer2DModel(in_channels=64, num_layers=2, num_single_layers=3, num_attention_heads=2, joint_attention_dim=96,
                                     pooled_projection_dim=48, guidance_embeds=guidance).init_synthetic(seed)


def _inputs(seed=0, hw=(8, 6), n_txt=19):
    g = torch.Generator().manual_seed(seed)
    n_img = hw[0] * hw[1]
    hs = torch.randn(1, n_img, 64, generator=g).bfloat16()
    enc = torch.randn(1, n_txt, 96, generator=g).bfloat16()
    pooled = torch.randn(1, 48, generator=g).bfloat16()
    img_ids, txt_ids = fr.make_ids(hw[0], hw[1], n_txt)
    return hs, enc, pooled, img_ids, txt_ids


@pytest.mark.parametrize("guidance,n_txt", [(True, 19), (False, 19), (True, 24)])
def test_flux_engine_single_forward_matches_oracle(emulated, guidance, n_txt):
    """n_txt = 24: V^T written by the projection GEMM into its column range; 19: the row-major V + transpose fallback."""
    model = _model(guidance)
    hs, enc, pooled, img_ids, txt_ids = _inputs(n_txt=n_txt)
    t, gd = torch.tensor([0.731]), (torch.tensor([3.5]) if guidance else None)
    ref_m = copy.deepcopy(model)
    ref_m.__class__ = type("RefFlux", (ref_m.__class__,), {})
    fr.install_magcache(type(ref_m), mc.tables()["flux_dev"], 28)
    m64 = copy.deepcopy(model).double()
    m64.__class__ = type("RefFlux64", (m64.__class__,), {})
    fr.install_magcache(type(m64), mc.tables()["flux_dev"], 28)
    ours = copy.deepcopy(model)
    ours.__class__ = type("OurFlux", (ours.__class__,), {})
    mc.init_magcache_flux(ours, 28)
    with torch.no_grad():
        ref = ref_m(hs, enc, pooled, t, img_ids, txt_ids, gd, return_dict=False)[0]
        with fr.exact():
            exact = m64(hs.double(), enc.double(), pooled.double(), t.double(), img_ids, txt_ids, None if gd is None else gd.double(),
                        return_dict=False)[0]
        out = ours(hs, enc, pooled, t, img_ids, txt_ids, gd)
    assert hasattr(out, "sample") and out.sample.shape == ref.shape == (1, 48, 64) and out.sample.dtype == torch.bfloat16
    e_ours, e_ref, e_vs = rel_l2(out.sample, exact), rel_l2(ref, exact), rel_l2(out.sample, ref)
    print(f"[flux emulated] ours vs fp64 {e_ours:.3e} | oracle(bf16) vs fp64 {e_ref:.3e} | ours vs oracle {e_vs:.3e}")
    assert e_ours <= 1.5 * e_ref + 1e-3
    assert e_vs <= 2.0 * e_ref + 1e-3
    # the cached residual is the image-stream delta
    assert rel_l2(ours.previous_residual[0], ref_m.previous_residual[0]) <= 2.0 * e_ref + 2e-2
    assert ours.cnt == ref_m.cnt == 1


def test_flux_engine_loop_hits_and_misses(emulated):
    """A whole 12-step generation: identical skip decisions (incl. the step-11 veto mapping), controller attributes, outputs."""
    model = _model(True, seed=1)
    hs, enc, pooled, img_ids, txt_ids = _inputs(1)
    steps = 12
    ref_m = copy.deepcopy(model)
    ref_m.__class__ = type("RefFluxL", (ref_m.__class__,), {})
    fr.install_magcache(type(ref_m), mc.tables()["flux_dev"], steps, thresh=0.24, K=5, retention_ratio=0.1)
    ours = copy.deepcopy(model)
    ours.__class__ = type("OurFluxL", (ours.__class__,), {})
    mc.init_magcache_flux(ours, steps, thresh=0.24, K=5, retention_ratio=0.1)
    assert type(ours).mag_ratios.tolist() == type(ref_m).mag_ratios.tolist()
    want = mc.MagCacheConfig("flux", 0.24, 5, 0.1, steps, table="flux_dev").schedule().tolist()
    skips = []
    with torch.no_grad():
        for i in range(steps + 2):
            t = torch.tensor([1.0 - (i % steps) / steps])
            x = hs * (1.0 - 0.03 * i)
            ref = ref_m(x, enc, pooled, t, img_ids, txt_ids, torch.tensor([3.5]), return_dict=False)[0]
            out = ours(x, enc, pooled, t, img_ids, txt_ids, torch.tensor([3.5]), return_dict=False)[0]
            skips.append(int(ref_m.last_skip))
            assert rel_l2(out, ref) <= 0.15, (i, rel_l2(out, ref))
            for attr in ("cnt", "accumulated_ratio", "accumulated_err", "accumulated_steps"):
                assert float(getattr(ours, attr)) == float(getattr(ref_m, attr)), (i, attr)
    assert skips[:steps] == want and 0 < sum(want) < steps


def test_flux_forward_refuses_unbuilt_side_paths(emulated):
    ours = _model()
    ours.__class__ = type("OurFluxX", (ours.__class__,), {})
    mc.init_magcache_flux(ours, 28)
    hs, enc, pooled, img_ids, txt_ids = _inputs()
    with pytest.raises(NotImplementedError):
        ours(hs, enc, pooled, torch.tensor([0.5]), img_ids, txt_ids, torch.tensor([3.5]), controlnet_block_samples=[hs])
    with pytest.raises(ValueError):
        ours(hs, enc, pooled, torch.tensor([0.5]), img_ids, txt_ids, None)  # guidance-distilled model without guidance


def test_flux_calibration_twin(emulated, capsys):
    """`magcache_flux_calibration` (magcache_flux.py:21-231): same outputs as the oracle twin, statistics from the second call on (finer than
    the reference's bf16-quantised ones: compared at 2e-2), lists printed on the last call and cleared at the wrap."""
    model = _model(True, seed=2)
    hs, enc, pooled, img_ids, txt_ids = _inputs(2)
    steps = 4
    ref_m = copy.deepcopy(model)
    ref_m.__class__ = type("RefFluxC", (ref_m.__class__,), {})
    type(ref_m).forward = fr.magcache_calibration
    type(ref_m).cnt, type(ref_m).num_steps = 0, steps
    type(ref_m).norm_ratio, type(ref_m).norm_std, type(ref_m).cos_dis, type(ref_m).previous_residual = [], [], [], None
    ours = copy.deepcopy(model)
    ours.__class__ = type("OurFluxC", (ours.__class__,), {})
    mc.init_magcache_flux_calibration(ours, steps)
    ratios_ref, ratios_ours = [], []
    with torch.no_grad():
        for i in range(steps):
            t = torch.tensor([1.0 - i / steps])
            x = hs * (1.0 - 0.1 * i)
            a = ref_m(x, enc, pooled, t, img_ids, txt_ids, torch.tensor([3.5]), return_dict=False)[0]
            if i < steps - 1:
                ratios_ref = list(ref_m.norm_ratio)
            b = ours(x, enc, pooled, t, img_ids, txt_ids, torch.tensor([3.5]), return_dict=False)[0]
            if i < steps - 1:
                ratios_ours = list(ours.norm_ratio)
            assert rel_l2(b, a) <= 0.15
    assert len(ratios_ref) == len(ratios_ours) == steps - 2
    for a, b in zip(ratios_ours, ratios_ref):
        assert abs(a - b) <= 2e-2 * abs(b), (ratios_ours, ratios_ref)
    assert ours.cnt == 0 and ours.norm_ratio == [] and "norm ratio" in capsys.readouterr().out
