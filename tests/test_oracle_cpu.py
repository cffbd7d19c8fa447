"""CPU tests of the oracle itself (it must be trustworthy before it checks anything) and of the repo layout rules."""
import copy
import json
import os
import re

import numpy as np
import pytest
import torch

from oracle import wan_ref
from oracle.controller_ref import calibration_stats

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_calibration_stats_against_reference_statements():
    """tests/golden/calib_stats.json was produced by executing magcache_generate.py:166-173 on these seeded tensors."""
    with open(os.path.join(ROOT, "tests", "golden", "calib_stats.json")) as f:
        cases = json.load(f)
    for c in cases:
        B, N, D = c["shape"]
        g = torch.Generator().manual_seed(c["seed"])
        r_prev = torch.randn(B, N, D, generator=g) * 0.1
        r_cur = r_prev * (0.97 + c["spread"] * torch.rand(B, N, 1, generator=g)) + 0.01 * torch.randn(B, N, D, generator=g)
        got = calibration_stats(r_cur, r_prev)
        assert got == (c["norm_ratio_raw"], c["norm_std_raw"], c["cos_dis_raw"])
        assert tuple(round(v, 5) for v in got) == (c["norm_ratio"], c["norm_std"], c["cos_dis"])


def _tiny():
    return wan_ref.WanModel(**wan_ref.CONFIGS["tiny"], text_dim=128, text_len=16).init_synthetic(0)


def test_oracle_forward_shapes_dtypes_and_skip_semantics():
    m = _tiny()
    m.__class__ = type("T", (m.__class__,), {})
    table = [1.0, 1.0] + [0.98] * 10
    wan_ref.install_magcache(m.__class__, table, 6, thresh=0.05, K=1, retention_ratio=0.2)
    lat = torch.randn(16, 2, 8, 8)
    ctx = torch.randn(9, 128)
    outs, skips = [], []
    with torch.no_grad():
        for i in range(12):
            o = m([lat], t=torch.tensor([900.0 - 50 * i]), context=[ctx], seq_len=2 * 4 * 4)
            assert o[0].shape == (16, 2, 8, 8) and o[0].dtype == torch.float32
            outs.append(o[0])
            skips.append(int(m.last_skip))
    # int(12*0.2)=2: calls 0,1 always compute; K=1 -> never two hits in a row per branch
    assert skips[:2] == [0, 0] and sum(skips) > 0
    for b in (0, 1):
        s = skips[b::2]
        assert all(not (s[i] and s[i + 1]) for i in range(len(s) - 1))
    assert m.cnt == 0 and m.accumulated_steps == [0, 0]  # wrapped and reset at cnt == num_steps
    assert m.residual_cache[0].dtype == torch.float32 and m.residual_cache[0].shape == (1, 32, 256)


def test_oracle_bf16_close_to_fp64_evaluation():
    m = _tiny()
    lat, ctx = torch.randn(16, 2, 8, 8), torch.randn(9, 128)
    a = copy.deepcopy(m)
    a.__class__ = type("A", (a.__class__,), {})
    wan_ref.install_magcache(a.__class__, [1.0] * 8, 4)
    b = copy.deepcopy(m).double()
    b.__class__ = type("B", (b.__class__,), {})
    wan_ref.install_magcache(b.__class__, [1.0] * 8, 4)
    with torch.no_grad():
        ya = a([lat], t=torch.tensor([400.0]), context=[ctx], seq_len=32)[0]
        with wan_ref.exact_fp64():
            yb = b([lat.double()], t=torch.tensor([400.0]), context=[ctx.double()], seq_len=32)[0]
    assert yb.dtype == torch.float64
    rel = float((ya.double() - yb).norm() / yb.norm())
    assert 0 < rel < 2e-2, rel


def test_oracle_i2v_branch_uses_image_tokens_and_y():
    """i2v restatement (magcache_generate.py:226-227, :233-234, :264-266 + upstream WanI2VCrossAttention / MLPProj): the output
    depends on `y` and on `clip_fea`, the bf16 emulation stays close to the fp64 evaluation, and the reference's assert fires
    when the extra inputs are missing."""
    m = wan_ref.WanModel(dim=256, ffn_dim=512, num_heads=2, num_layers=2, in_dim=36, text_dim=128, text_len=32, model_type="i2v",
                         clip_dim=64).init_synthetic(0)
    m.__class__ = type("I", (m.__class__,), {})
    wan_ref.install_magcache(m.__class__, [1.0] * 8, 4)
    b = copy.deepcopy(m).double()
    b.__class__ = type("I64", (b.__class__,), {})
    wan_ref.install_magcache(b.__class__, [1.0] * 8, 4)
    g = torch.Generator().manual_seed(0)
    lat, y, ctx, clip = torch.randn(16, 2, 8, 8, generator=g), torch.randn(20, 2, 8, 8, generator=g), torch.randn(9, 128, generator=g), torch.randn(1, 257, 64, generator=g)
    t = torch.tensor([400.0])
    with torch.no_grad():
        base = m([lat], t=t, context=[ctx], seq_len=32, clip_fea=clip, y=[y])[0]
        m.cnt = 0
        other_y = m([lat], t=t, context=[ctx], seq_len=32, clip_fea=clip, y=[y * 0.5])[0]
        m.cnt = 0
        other_clip = m([lat], t=t, context=[ctx], seq_len=32, clip_fea=clip.flip(1), y=[y])[0]
        with wan_ref.exact_fp64():
            exact = b([lat.double()], t=t, context=[ctx.double()], seq_len=32, clip_fea=clip.double(), y=[y.double()])[0]
        with pytest.raises(AssertionError):
            m([lat], t=t, context=[ctx], seq_len=32)
    assert base.shape == (16, 2, 8, 8) and base.dtype == torch.float32
    assert float((base - other_y).abs().max()) > 1e-3 and float((base - other_clip).abs().max()) > 1e-4
    rel = float((base.double() - exact).norm() / exact.norm())
    assert 0 < rel < 2e-2, rel


def test_oracle_vace_branch():
    """VACE restatement (magcache_generate.py:439-560 + upstream VaceWanModel): hints only on a miss, output depends on the control
    video and on vace_context_scale, scale 0 equals running the main blocks without hints, bf16 emulation close to fp64."""
    m = wan_ref.WanModel(dim=256, ffn_dim=512, num_heads=2, num_layers=4, text_dim=128, text_len=32, model_type="vace", vace_in_dim=24).init_synthetic(0)
    assert m.vace_layers == [0, 2] and [getattr(b, "block_id", None) for b in m.blocks] == [0, None, 1, None]
    assert hasattr(m.vace_blocks[0], "before_proj") and not hasattr(m.vace_blocks[1], "before_proj")
    m.__class__ = type("V", (m.__class__,), {})
    wan_ref.install_magcache(m.__class__, [1.0] * 8, 4, vace=True)
    g = torch.Generator().manual_seed(0)
    lat, vc, ctx = torch.randn(16, 2, 8, 8, generator=g), torch.randn(24, 2, 8, 8, generator=g), torch.randn(9, 128, generator=g)
    t = torch.tensor([300.0])

    def run(model, scale=1.0, vcx=vc, dt=torch.float32):
        model.cnt = 0
        return model([lat.to(dt)], t=t, vace_context=[vcx.to(dt)], context=[ctx.to(dt)], seq_len=32, vace_context_scale=scale)[0]

    with torch.no_grad():
        a, b, c0 = run(m), run(m, vcx=vc * 0.3), run(m, scale=0.0)
        plain = copy.deepcopy(m)
        plain.__class__ = type("P", (plain.__class__,), {})
        wan_ref.install_magcache(plain.__class__, [1.0] * 8, 4)  # the T2V forward on the same weights: no hints at all
        plain.cnt = 0
        p = plain([lat], t=t, context=[ctx], seq_len=32)[0]
        m64 = copy.deepcopy(m).double()
        m64.__class__ = type("V64", (m64.__class__,), {})
        wan_ref.install_magcache(m64.__class__, [1.0] * 8, 4, vace=True)
        with wan_ref.exact_fp64():
            exact = run(m64, dt=torch.float64)
    assert float((a - b).abs().max()) > 1e-2 and float((a - c0).abs().max()) > 1e-2
    assert torch.equal(c0, p)
    rel = float((a.double() - exact).norm() / exact.norm())
    assert 0 < rel < 2e-2, rel


def test_oracle_seq_len_padding_reaches_no_real_token():
    """magcache_generate.py:242-246 pads the token axis with zero rows up to `seq_len`; with the key mask of upstream's
    flash_attention(k_lens=seq_lens) the real tokens' outputs do not change (what lets the CUDA path skip the padded rows), while the
    cached residual gains seq_len - n rows."""
    m = _tiny()
    lat, ctx = torch.randn(16, 2, 8, 8), torch.randn(9, 128)
    outs, caches = [], []
    for seq_len in (32, 37):
        a = copy.deepcopy(m)
        a.__class__ = type("P", (a.__class__,), {})
        wan_ref.install_magcache(a.__class__, [1.0] * 8, 4)
        with torch.no_grad():
            outs.append(a([lat], t=torch.tensor([400.0]), context=[ctx], seq_len=seq_len)[0])
        caches.append(a.residual_cache[0])
    assert caches[0].shape[1] == 32 and caches[1].shape[1] == 37
    assert float((outs[0] - outs[1]).abs().max()) <= 1e-5 * float(outs[0].abs().max())  # SDPA picks another kernel with a mask
    assert float((caches[0][0] - caches[1][0, :32]).abs().max()) <= 2e-2 * float(caches[0].abs().max())


def test_denoise_loop_calls_cond_then_uncond():
    m = _tiny()
    m.__class__ = type("L", (m.__class__,), {})
    wan_ref.install_magcache(m.__class__, [1.0] * 100, 3, retention_ratio=0.5)  # int(6*0.5)=3: both cache slots filled first
    seen = []
    orig = m.__class__.forward

    def spy(self, x, t, context, seq_len, **kw):
        seen.append((self.cnt, float(context[0][0, 0])))
        return orig(self, x, t, context, seq_len, **kw)

    m.__class__.forward = spy
    c1, c0 = torch.full((4, 128), 1.0), torch.full((4, 128), -1.0)
    with torch.no_grad():
        out = wan_ref.denoise_loop(m, torch.randn(16, 1, 8, 8), c1, c0, steps=3)
    assert [s[1] for s in seen] == [1.0, -1.0] * 3 and [s[0] for s in seen] == [0, 1, 2, 3, 4, 5]
    assert out.shape == (16, 1, 8, 8) and torch.isfinite(out).all()


def test_product_never_imports_oracle_or_reference():
    """The product path must not route through the oracle (or read /root/reference) — enforced on the source text."""
    pkg = os.path.join(ROOT, "magcache_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), fn
                assert "/root/reference" not in src, fn
    for fn in ("bench.py", "__graft_entry__.py"):
        path = os.path.join(ROOT, fn)
        if os.path.exists(path):
            assert "/root/reference" not in open(path).read(), fn


def test_cabi_exports_every_declared_symbol():
    """include/magcache_b200.h and the shared library agree (no compute calls: works without a GPU)."""
    from magcache_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "magcache_b200.h")).read()
    declared = set(re.findall(r"\b(mc_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"mc_ctrl_config", "mc_ctrl_state"}
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(_lib.lib, name), f"{name} declared in the header but not exported"
    bound = set(_lib.SIGNATURES) | set(_lib.OTHER_EXPORTS)
    assert bound == declared, bound ^ declared
    assert _lib.lib.mc_abi_version() >= 1
    # error plumbing works without a device
    import ctypes
    assert _lib.lib.mc_cache_hit_add(None, 0, None, 0, None, 0, 8, None) == _lib.MC_ERR_INVALID
    assert b"null pointer" in _lib.lib.mc_last_error()
    assert _lib.lib.mc_gemm_bf16(ctypes.c_void_p(16), 8, ctypes.c_void_p(16), 8, 1, 1, 4, None, 0, ctypes.c_void_p(16), 8, None, None) == _lib.MC_ERR_INVALID


def test_ops_refuse_cpu_tensors():
    import pytest
    from magcache_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.cache_hit_add(torch.zeros(8, dtype=torch.bfloat16), torch.zeros(8))
