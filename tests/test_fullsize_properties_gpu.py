"""Size-independent properties checked at BASELINE.json's FULL sizes (Wan2.1-1.3B 832x480x81f: 32760 tokens x 1536), where a
CPU oracle run would take minutes per call: linearity of the tensor-core GEMM, attention's convex-combination invariants,
bit-equality of the fused cache-hit head with the unfused add + head, and run-to-run determinism of a full-shape block."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
N, D = 32760, 1536


def test_gemm_linearity_full_rows():
    """(A1 + A2) B^T == A1 B^T + A2 B^T up to fp32 accumulation order, for 32760 x 1536 x 1536 (ragged last M tile)."""
    from magcache_b200 import _lib, ops
    g = torch.Generator(device=DEV).manual_seed(0)
    # values on a coarse grid so that a1 + a2 is exact in bf16 and the identity holds up to accumulation order only
    a1 = (torch.randint(-8, 9, (N, D), device=DEV, generator=g).float() / 8).bfloat16()
    a2 = (torch.randint(-8, 9, (N, D), device=DEV, generator=g).float() / 8).bfloat16()
    b = (torch.randint(-8, 9, (D, D), device=DEV, generator=g).float() / 64).bfloat16()
    y1 = ops.gemm(a1, b, None, _lib.MC_EPI_BIAS_F32)
    y2 = ops.gemm(a2, b, None, _lib.MC_EPI_BIAS_F32)
    y12 = ops.gemm((a1.float() + a2.float()).bfloat16(), b, None, _lib.MC_EPI_BIAS_F32)
    assert torch.equal((a1.float() + a2.float()).bfloat16().float(), a1.float() + a2.float())
    assert torch.allclose(y12, y1 + y2, rtol=1e-3, atol=1e-4)  # the north-star tolerance
    assert torch.isfinite(y12).all() and float(y12.abs().max()) > 0


def test_attention_convexity_full_sequence():
    """softmax weights are a convex combination: with V == per-head constant rows the output equals that constant for every
    query (up to bf16 rounding of P and of the output), and the output is invariant to adding a constant to all scores."""
    from magcache_b200 import ops
    heads = 12
    g = torch.Generator(device=DEV).manual_seed(1)
    q = torch.randn(N, D, device=DEV, generator=g).bfloat16()
    k = torch.randn(N, D, device=DEV, generator=g).bfloat16()
    const = torch.linspace(-2, 2, D, device=DEV).bfloat16()
    v = const[None, :].expand(N, D).contiguous()
    out = ops.attention(q, k, v, heads)
    err = (out.float() - const.float()[None, :]).abs()
    assert float(err.max()) <= 2.0 ** -7 * 2 + 1e-3, float(err.max())
    # second invariant on a slice of queries: out depends on k only through q.k differences -> shifting every key by a vector
    # orthogonal to... is not exact in bf16; instead check the row-stochastic property with a one-hot V column block
    onehot = torch.zeros(N, D, dtype=torch.bfloat16, device=DEV)
    onehot[:N // 2] = 1.0
    out2 = ops.attention(q[:1024], k, onehot, heads).float()
    assert float(out2.min()) >= -1e-3 and float(out2.max()) <= 1.0 + 2.0 ** -7  # a probability mass in [0, 1]
    assert abs(float(out2.mean()) - 0.5) < 0.05  # iid keys: about half the mass on the first half of the keys


def test_fused_hit_equals_add_then_head_bitwise():
    """`x = x + residual_x` (magcache_generate.py:295) + head: the fused kernel (sum formed on the fly) and the unfused
    K1 add followed by the head kernel give bit-identical fp32 outputs at the full shape."""
    from magcache_b200 import ops
    g = torch.Generator(device=DEV).manual_seed(2)
    x0 = torch.randn(N, D, device=DEV, generator=g).bfloat16()
    r = torch.randn(N, D, device=DEV, generator=g) * 0.3
    head_mod = torch.randn(2, D, device=DEV, generator=g) / math.sqrt(D)
    e = torch.randn(1, D, device=DEV, generator=g) * 0.2
    wt = (torch.randn(D, 64, device=DEV, generator=g) * 0.03).contiguous()
    b = torch.randn(64, device=DEV, generator=g) * 0.1
    grid = (21, 30, 52)
    fused = ops.head_unpatchify(x0, head_mod, e, wt, b, grid, residual=r)
    unfused = ops.head_unpatchify(ops.cache_hit_add(x0, r), head_mod, e, wt, b, grid)
    assert fused.shape == (16, 21, 60, 104)
    assert torch.equal(fused, unfused)


def test_full_shape_block_is_deterministic_and_finite():
    """One Wan2.1-1.3B-shaped block at 32760 tokens through the engine twice: identical bits, finite, non-trivial; and a
    miss followed by a hit reproduces head(x0 + (x - x0)) to fp32 rounding of the residual round trip."""
    import magcache_b200 as mc
    dims = mc.WanDims(1536, 8960, 12, 1)
    model = mc.WanModelHandle(mc.WanWeights.random(dims, torch.device(DEV), seed=3))
    mc.init_magcache(model, 10, thresh=0.12, K=2, retention_ratio=0.1, mag_ratios=[1.0, 1.0] + [0.999] * 18)
    g = torch.Generator(device=DEV).manual_seed(4)
    lat = torch.randn(16, 21, 60, 104, device=DEV, generator=g)
    ctx = torch.randn(512, 4096, device=DEV, generator=g).bfloat16()
    t = torch.tensor([777.0], device=DEV)
    outs = []
    for _ in range(2):
        mc.reset_magcache(model)
        outs.append(model([lat], t=t, context=[ctx], seq_len=N)[0])
    assert torch.equal(outs[0], outs[1])
    assert torch.isfinite(outs[0]).all() and float(outs[0].abs().mean()) > 1e-3
    miss_uncond = model([lat], t=t, context=[ctx], seq_len=N)[0]   # cnt 1: miss (fills slot 1)
    hit = model([lat], t=t, context=[ctx], seq_len=N)[0]           # cnt 2: hit on slot 0 with the same inputs
    assert model.cnt == 3
    # same inputs => x0 + (x - x0) differs from x only by the fp32 rounding of the subtraction/addition pair
    rel = float((hit - outs[0]).norm() / outs[0].norm())
    assert rel < 1e-5, rel
    assert torch.equal(miss_uncond, outs[0])  # same inputs, same weights: the uncond-slot miss equals the cond-slot miss
