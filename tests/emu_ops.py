"""Test infrastructure: a torch-CPU emulation of magcache_b200.ops with the rounding points the kernels document (bf16 where the
kernel rounds to bf16, fp32 elsewhere). Monkeypatched over an engine module's `ops`, it lets the ORCHESTRATION of an engine — weight
packing, buffer views, row / column ranges, modulation indices, gates, op order — be checked against the oracle on CPU. It proves
nothing about the kernels themselves (tests/test_kernels_gpu.py and the *_gpu.py forward tests do that)."""
import math

import torch
import torch.nn.functional as F

from magcache_b200 import _lib as L

BF, F32 = torch.bfloat16, torch.float32
LAUNCHES = 0
PROFILE = None


def _al(t, n=16):
    """The kernels' pointer / pitch preconditions (mc_gemm_bf16, mc_attn_fwd, ... check them on the device side): CPU allocations are
    64-byte aligned like CUDA ones are 256-byte aligned, so a misaligned VIEW offset shows up here exactly as it would on the GPU."""
    return t.data_ptr() % n == 0


def _rb(t):
    return t.to(BF).to(F32)


def _count(n=1):
    global LAUNCHES
    LAUNCHES += n


def gemm(a, b, bias=None, epilogue=L.MC_EPI_BIAS_BF16, out=None, gate=None, tag=None):
    assert a.dtype == BF and b.dtype == BF and a.shape[1] == b.shape[1]
    assert a.stride(1) == 1 and b.stride(1) == 1 and a.shape[1] % 8 == 0 and a.stride(0) % 8 == 0 and b.stride(0) % 8 == 0, "TMA operand pitch"
    assert _al(a) and _al(b) and (out is None or _al(out)), "mc_gemm_bf16: A/B/out must be 16-byte aligned"
    acc = a.to(F32) @ b.to(F32).t()
    M, N = acc.shape
    if epilogue == L.MC_EPI_ROWBIAS_BF16:
        r = (acc + (bias[:, None] if bias is not None else 0.0)).to(BF)
    else:
        acc = acc + (bias[None, :] if bias is not None else 0.0)
        if epilogue == L.MC_EPI_BIAS_BF16:
            r = acc.to(BF)
        elif epilogue == L.MC_EPI_BIAS_GELU_BF16:
            r = F.gelu(_rb(acc), approximate="tanh").to(BF)
        elif epilogue == L.MC_EPI_BIAS_GELU_ERF_BF16:
            r = F.gelu(_rb(acc)).to(BF)
        elif epilogue == L.MC_EPI_BIAS_SILU_BF16:
            r = F.silu(_rb(acc)).to(BF)
        elif epilogue == L.MC_EPI_BIAS_F32:
            r = acc
        elif epilogue == L.MC_EPI_BIAS_GATE_RESID:
            assert out is not None and out.dtype == F32
            r = out + _rb(acc) * (gate[None, :] if gate is not None else 1.0)
        elif epilogue == L.MC_EPI_BIAS_GATE_RESID_BF16:
            assert out is not None and out.dtype == BF
            g = gate[None, :] if gate is not None else 1.0
            r = (out.to(F32) + _rb(g * _rb(acc))).to(BF)
        else:
            raise ValueError(epilogue)
    if out is None:
        out = torch.empty(M, N, dtype=r.dtype)
    assert out.shape == (M, N) and out.dtype == r.dtype and out.stride(1) == 1
    out.copy_(r)
    _count()
    return out


def ln_modulate(x, em, scale_idx, shift_idx, eps=1e-6, round_ln_to_bf16=False, out_dtype=BF, out=None):
    assert x.is_contiguous() and em.dtype == F32 and em.is_contiguous()
    ln = F.layer_norm(x.to(F32), (x.shape[-1],), None, None, eps)
    if round_ln_to_bf16:
        ln = _rb(ln)
    y = ln * (1.0 + em[scale_idx]) + em[shift_idx]
    if out is None:
        out = torch.empty(x.shape, dtype=out_dtype)
    assert out.is_contiguous()
    out.copy_(y)
    _count()
    return out


def ln_affine(x, weight, bias, eps=1e-6, out_dtype=BF, out=None):
    y = F.layer_norm(x.to(F32), (x.shape[-1],), weight, bias, eps)
    if out is None:
        out = torch.empty(x.shape, dtype=out_dtype)
    out.copy_(y)
    _count()
    return out


def rmsnorm_head_rope_(x, weight, heads, cos_sin=None, eps=1e-6):
    assert x.dtype == BF and x.stride(1) == 1 and x.shape[1] == heads * 128 and weight.numel() == 128
    assert x.stride(0) % 8 == 0 and _al(x) and (cos_sin is None or (cos_sin.is_contiguous() and _al(cos_sin, 32) and cos_sin.shape == (x.shape[0], 128)))
    rows = x.shape[0]
    v = x.to(F32).view(rows, heads, 128)
    r = torch.rsqrt(v.pow(2).mean(-1, keepdim=True) + eps)
    o = _rb(_rb(v * r) * weight)
    if cos_sin is not None:
        cs = cos_sin.view(rows, 1, 64, 2)
        re, im = o.view(rows, heads, 64, 2).unbind(-1)
        c, s = cs[..., 0], cs[..., 1]
        o = torch.stack([re * c - im * s, im * c + re * s], dim=-1).view(rows, heads, 128)
    x.copy_(o.reshape(rows, heads * 128))
    _count()
    return x


def rmsnorm_rope_segs_(x, weights, segs, cos_sin=None, head_dim=128, eps=1e-6, tag=None):
    cols = weights.shape[-1]
    assert weights.numel() == segs * cols and x.shape[1] >= segs * cols
    for sg in range(segs):
        rmsnorm_rope_(x[:, sg * cols:(sg + 1) * cols], weights.reshape(segs, cols)[sg], cos_sin, head_dim, eps)
    global LAUNCHES
    LAUNCHES -= segs - 1
    return x


def rmsnorm_rope_(x, weight, cos_sin=None, head_dim=128, eps=1e-6, tag=None):
    rows, cols = x.shape
    v = x.to(F32)
    o = _rb(v * torch.rsqrt(v.pow(2).mean(-1, keepdim=True) + eps)) * weight
    if cos_sin is not None:
        H = cols // head_dim
        cs = cos_sin.view(rows, 1, head_dim // 2, 2)
        re, im = o.view(rows, H, head_dim // 2, 2).unbind(-1)
        c, s = cs[..., 0], cs[..., 1]
        o = torch.stack([re * c - im * s, re * s + im * c], dim=-1).view(rows, cols)
    x.copy_(o)
    _count()
    return x


def attention(q, k, v, heads, scale=None, out=None, tag=None, first_key_row=0, seg_flags=None, seg_epoch=None, seg_rows=0):
    Lq, W = q.shape
    Lk = k.shape[0]
    assert q.stride(1) == 1 and k.stride(1) == 1 and v.stride(1) == 1 and v.shape == (Lk, W) and W == heads * 128
    assert q.stride(0) % 8 == 0 and k.stride(0) % 8 == 0 and v.stride(0) % 8 == 0 and _al(q) and _al(k) and _al(v), "attention TMA operands"
    assert out is None or (out.stride(1) == 1 and out.stride(0) % 8 == 0 and _al(out))
    qh = q.to(F32).reshape(Lq, heads, 128).transpose(0, 1)
    kh = k.to(F32).reshape(Lk, heads, 128).transpose(0, 1)
    vh = v.to(F32).reshape(Lk, heads, 128).transpose(0, 1)
    s = qh @ kh.transpose(1, 2) * (scale if scale is not None else 1.0 / math.sqrt(128))
    o = (torch.softmax(s, -1) @ vh).transpose(0, 1).reshape(Lq, W)
    if out is None:
        out = torch.empty(Lq, W, dtype=BF)
    out.copy_(o)
    _count()
    return out


def _promote(a, b):
    return F32 if F32 in (a.dtype, b.dtype) else BF


def cache_hit_add(x, r, out=None, tag=None):
    assert x.shape == r.shape and x.is_contiguous() and r.is_contiguous() and (out is None or out.is_contiguous())
    y = x.to(F32) + r.to(F32)
    if out is None:
        out = torch.empty(x.shape, dtype=_promote(x, r))
    out.copy_(y)
    _count()
    return out


def residual_sub(x_out, x_in, out=None):
    assert x_out.shape == x_in.shape and x_out.is_contiguous() and x_in.is_contiguous() and (out is None or out.is_contiguous())
    y = x_out.to(F32) - x_in.to(F32)
    if out is None:
        out = torch.empty(x_out.shape, dtype=_promote(x_out, x_in))
    out.copy_(y)
    _count()
    return out


def cast_into(src, dst):
    assert src.is_contiguous() and dst.is_contiguous() and src.numel() == dst.numel()
    dst.copy_(src.reshape(dst.shape))
    _count()
    return dst


def colmean(x):
    _count()
    return (x.to(F32).sum(0, keepdim=True).to(BF).to(F32) / torch.tensor(float(x.shape[0])).to(BF).to(F32)).to(BF)


def silu(x, out=None):
    y = F.silu(x.to(F32)).to(BF)
    if out is None:
        return y
    out.copy_(y)
    return out


def time_sinusoid(t, dim):
    half = dim // 2
    pos = t.to(torch.float64).reshape(-1)
    s = torch.outer(pos, torch.pow(10000.0, -torch.arange(half, dtype=torch.float64) / half))
    _count()
    return torch.cat([torch.cos(s), torch.sin(s)], dim=1).to(F32)


def transpose(src, dst):
    assert src.dtype == BF and dst.dtype == BF and src.stride(1) == 1 and dst.stride(1) == 1 and dst.shape == (src.shape[1], src.shape[0])
    dst.copy_(src.t())
    _count()
    return dst


def patchify(latent):
    C, Fr, H, W = latent.shape
    x = latent.view(C, Fr, H // 2, 2, W // 2, 2).permute(1, 2, 4, 0, 3, 5).reshape(Fr * (H // 2) * (W // 2), C * 4)
    _count()
    return x.to(BF)


def linear_f32_small(x, w, b=None, act=0):
    xin = F.silu(x) if act == 1 else x
    y = xin @ w.t() + (b if b is not None else 0.0)
    _count()
    return F.silu(y) if act == 2 else y


def head_prepare(head_mod, e, w_t, b, tag=None, slot=0):
    assert w_t.shape == (head_mod.shape[-1], 64) and w_t.is_contiguous() and b.numel() == 64 and e.numel() == head_mod.shape[-1]
    _count()
    return None


def head_unpatchify(x, head_mod, e, w_t, b, grid, c_out=16, residual=None, eps=1e-6, tag=None, row_offset=0, out=None, round_sum_to_bf16=False,
                    peer_outs=None, prep=None, step=None):
    assert (x.dtype == F32 and residual is None) or (x.dtype == BF and residual is not None), "fp32 stream, or bf16 + fp32 residual"
    assert x.shape[1] % 64 == 0, "mc_head_unpatchify: cols % 64"
    xs = x.to(F32) + (residual if residual is not None else 0.0)
    if round_sum_to_bf16:
        xs = _rb(xs)
    em = head_mod + e.reshape(1, -1)  # (modulation[1,2,D] + e.unsqueeze(1)).chunk(2): shift, scale
    y = F.layer_norm(xs, (xs.shape[-1],), None, None, eps) * (1 + em[1]) + em[0]
    o = y @ w_t + b  # [rows, 4*c_out]
    f, h, w = grid
    assert c_out == 16 and o.shape[1] == 64, "the head kernel writes 16 channels (64 features) per launch"
    unpatch = lambda rows: torch.einsum("fhwpqrc->cfphqwr", rows.view(f, h, w, 1, 2, 2, c_out)).reshape(c_out, f, 2 * h, 2 * w).contiguous()  # noqa: E731
    written = None
    if o.shape[0] != f * h * w:  # a row range [row_offset, row_offset + rows): only the positions of those rows are written
        assert out is not None, "a partial token range needs a caller-provided output"
        full, mark = torch.zeros(f * h * w, o.shape[1]), torch.zeros(f * h * w, o.shape[1])
        full[row_offset:row_offset + o.shape[0]] = o
        mark[row_offset:row_offset + o.shape[0]] = 1.0
        o, written = full, unpatch(mark) > 0
    u = unpatch(o)
    if step is not None:  # mc_head_unpatchify_step: the caller loop's CFG combine + scheduler update in the epilogue (full token range)
        assert o.shape[0] == f * h * w
        cond, x_lat, g, cx, cv = step
        t32 = lambda a: torch.tensor(a, dtype=F32)  # noqa: E731
        v = u + t32(g) * (cond - u)
        u = t32(cx) * x_lat + t32(cv) * v
    _count()
    if out is not None:
        assert out.is_contiguous() and out.shape == u.shape
        if written is None:
            out.copy_(u)
        else:
            out[written] = u[written]
        return out
    return u


def cfg_combine(cond, uncond, guide_scale, out=None):
    r = uncond + torch.tensor(guide_scale, dtype=F32) * (cond - uncond)
    _count()
    if out is None:
        return r
    out.copy_(r)
    return out


def cfg_step(cond, uncond, guide_scale, x, coef_v, coef_x=1.0, hist=(), coef_h=(), sigma=0.0, out=None, x0_out=None):
    f = lambda a: torch.tensor(a, dtype=F32)  # noqa: E731
    v = uncond + f(guide_scale) * (cond - uncond)
    acc = f(coef_x) * x + f(coef_v) * v
    for h, c in zip(hist, coef_h):
        acc = acc + f(c) * h
    if x0_out is not None:
        x0_out.copy_(x - f(sigma) * v)
    _count()
    if out is None:
        return acc
    out.copy_(acc)
    return out


def rel_l1(cur, prev):
    _count()
    return ((cur - prev).abs().mean() / prev.abs().mean()).item()


def residual_stats(r_cur, r_prev, denom_eps=0.0, reduce=None):
    if reduce is None:
        from oracle.controller_ref import calibration_stats
        _count(2)
        return calibration_stats(r_cur.to(F32)[None], r_prev.to(F32)[None], denom_eps)
    # token-sharded caller: the kernel's four raw sums over this rank's rows, reduced, then finalised like ops._finish_stats
    zero = torch.zeros_like(r_cur, dtype=BF)
    return residual_sub_stats(r_cur.to(F32), zero, r_prev, denom_eps, reduce)[1]


def residual_sub_stats(x_out, x_in, r_prev, denom_eps=0.0, reduce=None):
    """`mc_residual_sub_stats` as the kernel defines it: r = x_out - x_in (fp32) and the four raw sums (sum ratio, sum ratio^2,
    sum (1 - cos), rows) over the rows, handed to `reduce` before they are finalised like ops._finish_stats does."""
    import math

    import torch.nn.functional as Fn
    r = x_out.to(F32) - x_in.to(F32)
    ratio = (r.norm(dim=-1) / (r_prev.to(F32).norm(dim=-1) + denom_eps)).double()
    cosd = (1 - Fn.cosine_similarity(r, r_prev.to(F32), dim=-1, eps=1e-8)).double()
    stats = torch.stack([ratio.sum(), (ratio * ratio).sum(), cosd.sum(), torch.tensor(float(r.shape[0]), dtype=torch.float64)])
    _count(2)
    if reduce is not None:
        stats = reduce(stats)
    s0, s1, s2, n = stats.tolist()
    mean = s0 / n
    var = (s1 - s0 * s0 / n) / (n - 1) if n > 1 else float("nan")
    return r, (mean, math.sqrt(max(var, 0.0)), s2 / n)
