"""FLUX MMDiT engine on the GPU: the new kernels (per-head RMSNorm + RoPE, SiLU, the bf16 gated-residual and SiLU GEMM epilogues) against
torch, and `magcache_flux_forward` against the oracle restatement of MagCache4FLUX/magcache_flux.py:234-440.

Tolerances are tied to the oracle's own error: every comparison against the bf16 oracle is bounded by a small multiple of the oracle's
distance from an fp64 evaluation of the same network on the same inputs (all-bf16 streams with synthetic weights amplify rounding to
several percent, so a fixed number would be either meaningless or flaky)."""
import copy
import math
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def _ops():
    from magcache_b200 import ops
    return ops


def test_rmsnorm_head_rope_vs_torch():
    import emu_ops
    ops = _ops()
    for rows, heads, strided in [(300, 24, False), (77, 2, True), (4608, 24, True)]:
        W = heads * 128
        buf = torch.randn(rows, 2 * W, device=DEV).bfloat16()
        x = buf[:, W:] if strided else buf[:, :W].contiguous()
        w = (1 + 0.1 * torch.randn(128, device=DEV)).bfloat16().float()
        ang = torch.rand(rows, 64, device=DEV, dtype=torch.float64) * 6.28
        cs = torch.stack([ang.cos(), ang.sin()], dim=-1).reshape(rows, 128).float().contiguous()
        for table in (cs, None):
            want = x.clone().cpu()
            emu_ops.rmsnorm_head_rope_(want, w.cpu(), heads, None if table is None else table.cpu())
            got = x.clone()
            ops.rmsnorm_head_rope_(got, w, heads, table)
            d = (got.float().cpu() - want.float()).abs()
            # rsqrtf vs torch.rsqrt can flip the bf16 rounding of a normalised value by one ulp (2^-8 relative); through the rotation
            # that is at most 2^-8 (|re| + |im|) + the output rounding < 2^-7 * |(re, im)| — bounded by the PAIR norm, not by the
            # (possibly cancelled) output element. Twice that as the hard limit, and such flips must stay rare.
            pair = want.float().view(rows, heads, 64, 2).norm(dim=-1, keepdim=True).expand(rows, heads, 64, 2).reshape(rows, W)
            assert bool((d <= 2.0 ** -6 * pair + 1e-6).all()), float((d / (pair + 1e-6)).max())
            assert float((d > 0).float().mean()) < 0.02


def test_silu_and_gemm_epilogues_6_7_vs_torch():
    import emu_ops
    from magcache_b200 import _lib as L
    ops = _ops()
    x = torch.randn(1, 3072, device=DEV).bfloat16()
    assert torch.equal(ops.silu(x).cpu(), emu_ops.silu(x.cpu()))
    for M, N, K in [(391, 640, 1536), (1, 768, 256), (4608, 3072, 15360 // 4)]:
        a = torch.randn(M, K, device=DEV).bfloat16()
        b = (torch.randn(N, K, device=DEV) / math.sqrt(K)).bfloat16()
        bias = torch.randn(N, device=DEV).bfloat16().float()
        gate = (torch.randn(N, device=DEV) * 0.5).bfloat16().float()
        out = ops.gemm(a, b, bias, L.MC_EPI_BIAS_SILU_BF16)
        ref = emu_ops.gemm(a.cpu(), b.cpu(), bias.cpu(), L.MC_EPI_BIAS_SILU_BF16)
        acc = (a.double() @ b.double().t() + bias.double()).float().cpu()
        assert ((out.float().cpu() - ref.float()).abs() <= ref.float().abs() * 2.0 ** -7 + acc.abs() * 2.0 ** -7 + 1e-3).all()
        stream = torch.randn(M, N, device=DEV).bfloat16()
        want = emu_ops.gemm(a.cpu(), b.cpu(), bias.cpu(), L.MC_EPI_BIAS_GATE_RESID_BF16, out=stream.clone().cpu(), gate=gate.cpu())
        got = ops.gemm(a, b, bias, L.MC_EPI_BIAS_GATE_RESID_BF16, out=stream.clone(), gate=gate)
        tol = want.float().abs() * 2.0 ** -7 + (gate.cpu().abs() * acc.abs()) * 2.0 ** -6 + 1e-3
        assert ((got.float().cpu() - want.float()).abs() <= tol).all()


def _model(guidance=True, seed=0, layers=(2, 3)):
    from oracle import flux_ref as fr
    return fr.FluxTransformer2DModel(in_channels=64, num_layers=layers[0], num_single_layers=layers[1], num_attention_heads=2,
                                     joint_attention_dim=96, pooled_projection_dim=48, guidance_embeds=guidance).init_synthetic(seed)


def test_flux_forward_vs_oracle_and_fp64():
    import magcache_b200 as mc
    from oracle import flux_ref as fr
    model = _model()
    g = torch.Generator().manual_seed(0)
    hs, enc, pooled = torch.randn(1, 16 * 12, 64, generator=g).bfloat16(), torch.randn(1, 40, 96, generator=g).bfloat16(), torch.randn(1, 48, generator=g).bfloat16()
    img_ids, txt_ids = fr.make_ids(16, 12, 40)
    t, gd = torch.tensor([0.731]), torch.tensor([3.5])
    ref_m = copy.deepcopy(model)
    ref_m.__class__ = type("RefFlux", (ref_m.__class__,), {})
    fr.install_magcache(type(ref_m), mc.tables()["flux_dev"], 28)
    m64 = copy.deepcopy(model).double()
    m64.__class__ = type("RefFlux64", (m64.__class__,), {})
    fr.install_magcache(type(m64), mc.tables()["flux_dev"], 28)
    ours = copy.deepcopy(model).to(DEV)
    ours.__class__ = type("OurFlux", (ours.__class__,), {})
    mc.init_magcache_flux(ours, 28)
    with torch.no_grad():
        ref = ref_m(hs, enc, pooled, t, img_ids, txt_ids, gd, return_dict=False)[0]
        with fr.exact():
            exact = m64(hs.double(), enc.double(), pooled.double(), t.double(), img_ids, txt_ids, gd.double(), return_dict=False)[0]
    out = ours(hs.to(DEV), enc.to(DEV), pooled.to(DEV), t.to(DEV), img_ids.to(DEV), txt_ids.to(DEV), gd.to(DEV)).sample.cpu()
    e_ours, e_ref, e_vs = rel_l2(out, exact), rel_l2(ref, exact), rel_l2(out, ref)
    print(f"[flux] ours vs fp64 {e_ours:.3e} | oracle(bf16) vs fp64 {e_ref:.3e} | ours vs oracle {e_vs:.3e}")
    assert e_ours <= 1.5 * e_ref + 1e-3 and e_vs <= 2.0 * e_ref + 1e-3


def test_flux_loop_vs_oracle():
    """14 calls (hits and misses) through both patched forwards: controller state bit-equal on every call; tensors within twice the
    bf16 oracle's own distance from an fp64 run of the same loop (+1e-3)."""
    import magcache_b200 as mc
    from oracle import flux_ref as fr
    model = _model(seed=1)
    g = torch.Generator().manual_seed(1)
    hs, enc, pooled = torch.randn(1, 8 * 8, 64, generator=g).bfloat16(), torch.randn(1, 19, 96, generator=g).bfloat16(), torch.randn(1, 48, generator=g).bfloat16()
    img_ids, txt_ids = fr.make_ids(8, 8, 19)
    steps = 12
    ref_m = copy.deepcopy(model)
    ref_m.__class__ = type("RefFluxL", (ref_m.__class__,), {})
    fr.install_magcache(type(ref_m), mc.tables()["flux_dev"], steps)
    m64 = copy.deepcopy(model).double()
    m64.__class__ = type("RefFluxL64", (m64.__class__,), {})
    fr.install_magcache(type(m64), mc.tables()["flux_dev"], steps)
    ours = copy.deepcopy(model).to(DEV)
    ours.__class__ = type("OurFluxL", (ours.__class__,), {})
    mc.init_magcache_flux(ours, steps)
    skips = []
    with torch.no_grad():
        for i in range(steps + 2):
            t = torch.tensor([1.0 - (i % steps) / steps])
            x = hs * (1.0 - 0.03 * i)
            ref = ref_m(x, enc, pooled, t, img_ids, txt_ids, torch.tensor([3.5]), return_dict=False)[0]
            with fr.exact():
                exact = m64(x.double(), enc.double(), pooled.double(), t.double(), img_ids, txt_ids, torch.tensor([3.5]).double(), return_dict=False)[0]
            out = ours(x.to(DEV), enc.to(DEV), pooled.to(DEV), t.to(DEV), img_ids.to(DEV), txt_ids.to(DEV), torch.tensor([3.5], device=DEV),
                       return_dict=False)[0].cpu()
            skips.append(int(ref_m.last_skip))
            e_ref, e_vs, e_ours = rel_l2(ref, exact), rel_l2(out, ref), rel_l2(out, exact)
            assert e_vs <= 2.0 * e_ref + 1e-3, (i, e_vs, e_ref)
            assert e_ours <= 1.5 * e_ref + 1e-3, (i, e_ours, e_ref)
            for attr in ("cnt", "accumulated_ratio", "accumulated_err", "accumulated_steps"):
                assert float(getattr(ours, attr)) == float(getattr(ref_m, attr)), (i, attr)
    assert 0 < sum(skips[:steps]) < steps


def test_flux_mid_size_forward_vs_oracle_and_fp64():
    """A FLUX-shaped model at a mid size — hidden 1536 (12 heads x 128), 3 double + 5 single blocks, 32 x 32 = 1024 image tokens + 128
    text tokens (the long attention kernel, multi-tile GEMMs, ragged last tiles) — miss then hit, against the bf16 oracle and the
    fp64 evaluation."""
    import magcache_b200 as mc
    from oracle import flux_ref as fr
    model = fr.FluxTransformer2DModel(in_channels=64, num_layers=3, num_single_layers=5, num_attention_heads=12, joint_attention_dim=256,
                                      pooled_projection_dim=96, guidance_embeds=True).init_synthetic(4)
    g = torch.Generator().manual_seed(4)
    hs, enc, pooled = torch.randn(1, 1024, 64, generator=g).bfloat16(), torch.randn(1, 128, 256, generator=g).bfloat16(), torch.randn(1, 96, generator=g).bfloat16()
    img_ids, txt_ids = fr.make_ids(32, 32, 128)
    t, gd = torch.tensor([0.62]), torch.tensor([3.5])
    steps = 4
    table = [1.0] + [0.98] * 3
    ref_m = copy.deepcopy(model)
    ref_m.__class__ = type("RefFluxM", (ref_m.__class__,), {})
    fr.install_magcache(type(ref_m), table, steps, thresh=10.0, K=3, retention_ratio=0.25)
    m64 = copy.deepcopy(model).double()
    m64.__class__ = type("RefFluxM64", (m64.__class__,), {})
    fr.install_magcache(type(m64), table, steps, thresh=10.0, K=3, retention_ratio=0.25)
    ours = copy.deepcopy(model).to(DEV)
    ours.__class__ = type("OurFluxM", (ours.__class__,), {})
    mc.init_magcache_flux(ours, steps, thresh=10.0, K=3, retention_ratio=0.25, mag_ratios=table)
    kinds = []
    with torch.no_grad():
        for call in range(3):
            ref = ref_m(hs, enc, pooled, t, img_ids, txt_ids, gd, return_dict=False)[0]
            with fr.exact():
                exact = m64(hs.double(), enc.double(), pooled.double(), t.double(), img_ids, txt_ids, gd.double(), return_dict=False)[0]
            out = ours(hs.to(DEV), enc.to(DEV), pooled.to(DEV), t.to(DEV), img_ids.to(DEV), txt_ids.to(DEV), gd.to(DEV), return_dict=False)[0].cpu()
            kinds.append(int(ref_m.last_skip))
            e_ours, e_ref, e_vs = rel_l2(out, exact), rel_l2(ref, exact), rel_l2(out, ref)
            print(f"[flux mid, call {call}, {'hit' if kinds[-1] else 'miss'}] ours vs fp64 {e_ours:.3e} | oracle(bf16) vs fp64 {e_ref:.3e} | ours vs oracle {e_vs:.3e}")
            assert e_ours <= 1.5 * e_ref + 1e-3 and e_vs <= 2.0 * e_ref + 1e-3, (call, e_ours, e_ref, e_vs)
            assert float(ours.cnt) == float(ref_m.cnt)
    assert kinds == [0, 1, 1], kinds
