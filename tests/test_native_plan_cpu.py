"""`mc_dit_forward` (csrc/dit_forward.cu) against the Python engine WITHOUT a GPU: `mc_dit_plan` writes the launch plan of
the native forward — one line per launch, every operand by name and byte offset — and this test records the sequence `WanEngine` issues
through its ops module (kernels emulated, tests/emu_ops.py) in the same notation. The two must be identical line by line, for the
miss and the hit branch: same entry points, same order, same operands, same leading dimensions, same epilogues. (That the two paths
then produce the same bits is checked on the GPU, tests/test_native_forward_gpu.py.)"""
import ctypes
import types

import pytest
import torch

import magcache_b200 as mc
from magcache_b200 import _lib, native
from magcache_b200 import patch as patch_mod
from magcache_b200 import wan as wan_mod
from oracle import wan_ref

import emu_ops

DT = {torch.float32: 0, torch.bfloat16: 1}


class Recorder:
    """Stands in for the engine's `ops` module: forwards every call to the emulation and writes the launch in mc_dit_plan's notation."""

    def __init__(self):
        self.lines, self.ranges, self._small = [], [], iter(("e_h", "e", "e0"))

    def reg(self, t, name):
        self.ranges.append((t.data_ptr(), t.numel() * t.element_size(), name))

    def nm(self, t):
        if t is None:
            return "null"
        p = t.data_ptr()
        for base, nbytes, name in self.ranges:
            if base <= p < base + nbytes:
                return f"{name}+{p - base}"
        return "?"

    def __getattr__(self, name):  # everything not recorded explicitly must not be called by the plain forward
        raise AssertionError(f"the plain forward called ops.{name}, which mc_dit_forward does not sequence")

    def _count(self, n=1):
        pass

    PROFILE = None

    def patchify(self, latent):
        out = emu_ops.patchify(latent)
        self.reg(out, "tok")
        C, F, H, W = latent.shape
        self.lines.append(f"patchify {self.nm(latent)} C={C} F={F} H={H} W={W} -> {self.nm(out)}")
        return out

    def gemm(self, a, b, bias=None, epilogue=_lib.MC_EPI_BIAS_BF16, out=None, gate=None, tag=None):
        assert out is not None
        self.lines.append(f"gemm A={self.nm(a)} lda={a.stride(0)} B={self.nm(b)} ldb={b.stride(0)} M={a.shape[0]} N={b.shape[0]} K={a.shape[1]} "
                          f"bias={self.nm(bias)} epi={epilogue} out={self.nm(out)} ldo={out.stride(0)} gate={self.nm(gate)}")
        return emu_ops.gemm(a, b, bias, epilogue, out=out, gate=gate)

    def time_sinusoid(self, t, dim):
        out = emu_ops.time_sinusoid(t, dim)
        self.reg(out, "sin")
        self.lines.append(f"time_sinusoid {self.nm(t)} n={t.numel()} dim={dim} -> {self.nm(out)}")
        return out

    def linear_f32_small(self, x, w, b=None, act=0):
        y = emu_ops.linear_f32_small(x, w, b, act)
        self.reg(y, next(self._small))
        self.lines.append(f"linear_f32_small x={self.nm(x)} M={x.shape[0]} K={x.shape[1]} W={self.nm(w)} b={self.nm(b)} N={w.shape[0]} act={act} -> {self.nm(y)}")
        return y

    def head_prepare(self, head_mod, e, w_t, b, tag=None, slot=0):
        assert slot == 0
        self.lines.append(f"head_prepare mod={self.nm(head_mod)} e={self.nm(e)} Wt={self.nm(w_t)} b={self.nm(b)} cols={w_t.shape[0]} -> head_prep+0")
        return emu_ops.head_prepare(head_mod, e, w_t, b)

    def cache_hit_add(self, x, r, out=None, tag=None):
        assert out is not None
        self.lines.append(f"add x={self.nm(x)}:{DT[x.dtype]} r={self.nm(r)}:{DT[r.dtype]} -> {self.nm(out)}:{DT[out.dtype]} n={x.numel()}")
        return emu_ops.cache_hit_add(x, r, out=out)

    def residual_sub(self, x_out, x_in, out=None, tag=None):
        assert out is not None and x_out.dtype == torch.float32 and x_in.dtype == torch.bfloat16 and out.dtype == torch.float32
        self.lines.append(f"residual_sub x_out={self.nm(x_out)} x_in={self.nm(x_in)} -> {self.nm(out)} n={x_out.numel()}")
        return emu_ops.residual_sub(x_out, x_in, out=out)

    def cast_into(self, src, dst, tag=None):
        self.lines.append(f"cast {self.nm(src)}:{DT[src.dtype]} -> {self.nm(dst)}:{DT[dst.dtype]} n={src.numel()}")
        return emu_ops.cast_into(src, dst)

    def ln_modulate(self, x, em, scale_idx, shift_idx, eps=1e-6, round_ln_to_bf16=False, out_dtype=torch.bfloat16, out=None, tag=None):
        assert x.dtype == torch.float32 and out is not None and out.dtype == torch.bfloat16
        self.lines.append(f"ln_modulate x={self.nm(x)} rows={x.shape[0]} cols={x.shape[1]} mode=0 p0={self.nm(em)} p1=null scale={scale_idx} "
                          f"shift={shift_idx} round={int(round_ln_to_bf16)} -> {self.nm(out)}")
        return emu_ops.ln_modulate(x, em, scale_idx, shift_idx, eps=eps, round_ln_to_bf16=round_ln_to_bf16, out=out)

    def ln_affine(self, x, weight, bias, eps=1e-6, out_dtype=torch.bfloat16, out=None, tag=None):
        assert x.dtype == torch.float32 and out is not None and out.dtype == torch.bfloat16
        self.lines.append(f"ln_modulate x={self.nm(x)} rows={x.shape[0]} cols={x.shape[1]} mode=1 p0={self.nm(weight)} p1={self.nm(bias)} scale=0 "
                          f"shift=0 round=0 -> {self.nm(out)}")
        return emu_ops.ln_affine(x, weight, bias, eps=eps, out=out)

    def rmsnorm_rope_segs_(self, x, weights, segs, cos_sin=None, head_dim=128, eps=1e-6, tag=None):
        assert head_dim == 128
        self.lines.append(f"rmsnorm_rope_segs x={self.nm(x)} ld={x.stride(0)} rows={x.shape[0]} segs={segs} cols={weights.shape[-1]} w={self.nm(weights)} "
                          f"rope={self.nm(cos_sin)}")
        return emu_ops.rmsnorm_rope_segs_(x, weights, segs, cos_sin, head_dim, eps)

    def rmsnorm_rope_(self, x, weight, cos_sin=None, head_dim=128, eps=1e-6, tag=None):
        assert head_dim == 128
        self.lines.append(f"rmsnorm_rope x={self.nm(x)} ld={x.stride(0)} rows={x.shape[0]} cols={x.shape[1]} w={self.nm(weight)} rope={self.nm(cos_sin)}")
        return emu_ops.rmsnorm_rope_(x, weight, cos_sin, head_dim, eps)

    def attention(self, q, k, v, heads, scale=None, out=None, tag=None, **kw):
        assert not kw and scale is None and out is not None
        self.lines.append(f"attention q={self.nm(q)} ldq={q.stride(0)} k={self.nm(k)} ldk={k.stride(0)} v={self.nm(v)} ldv={v.stride(0)} "
                          f"out={self.nm(out)} Lq={q.shape[0]} Lk={k.shape[0]} heads={heads}")
        return emu_ops.attention(q, k, v, heads, out=out)

    def head_unpatchify(self, x, head_mod, e, w_t, b, grid, c_out=16, residual=None, eps=1e-6, tag=None, row_offset=0, out=None,
                        round_sum_to_bf16=False, peer_outs=None, prep=None, step=None):
        assert row_offset == 0 and out is None and peer_outs is None and step is None and not round_sum_to_bf16 and c_out == 16
        self.lines.append(f"head x={self.nm(x)}:{DT[x.dtype]} r={self.nm(residual)} rows={x.shape[0]} cols={x.shape[1]} prep=head_prep+0 -> out+0")
        return emu_ops.head_unpatchify(x, head_mod, e, w_t, b, grid, c_out=c_out, residual=residual, eps=eps)


def _engine_lines(eng, kind, slot):
    """Run one forward of the Python engine with the recorder as its ops module; returns the recorded plan."""
    rec = Recorder()
    w = eng.w
    for name in ("x0", "xs", "h", "att", "ffn", "cq", "qkv", "ckv", "ctx_h", "ctx", "em"):
        rec.reg(getattr(eng, name), name)
    rec.reg(eng.s_lat, "latent"), rec.reg(eng.s_t, "t"), rec.reg(eng.ctx_in, "context"), rec.reg(eng.res[slot], "residual")
    rec.reg(eng._rope_for(eng.grid), "rope")
    for name in _lib.DIT_TOP_FIELDS:
        rec.reg(getattr(w, name), name)
    for i, b in enumerate(w.blocks):
        for name in _lib.DIT_BLOCK_FIELDS:
            rec.reg(b[name], f"blk{i}.{name}")
    old = wan_mod.ops
    wan_mod.ops = rec
    try:
        eng.forward(kind, slot)
    finally:
        wan_mod.ops = old
    return rec.lines


@pytest.fixture()
def emulated(monkeypatch):
    monkeypatch.setattr(wan_mod, "ops", emu_ops)
    monkeypatch.setattr(patch_mod, "ops", emu_ops)
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True))


@pytest.mark.parametrize("cfg", [dict(dim=256, ffn_dim=512, num_heads=2, num_layers=2), dict(dim=384, ffn_dim=1024, num_heads=3, num_layers=3)])
def test_native_plan_equals_the_python_engines_sequence(emulated, cfg):
    model = wan_ref.WanModel(**cfg, text_dim=128, text_len=32).init_synthetic(2)
    weights = mc.WanWeights.from_module(model, torch.device("cpu"))
    eng = mc.WanEngine(weights)
    g = torch.Generator().manual_seed(0)
    lat, ctx = torch.randn(16, 3, 8, 12, generator=g), torch.randn(9, 128, generator=g)
    eng.stage_inputs(lat, torch.tensor([640.0]), ctx)
    nat = native.NativeWanForward(weights)
    need = nat.workspace_bytes(eng.grid)
    n_tok = 3 * 4 * 6
    assert need >= n_tok * (cfg["dim"] * (2 * 4 + 4 + 3 * 2) + cfg["ffn_dim"] * 2)  # at least x0, h, att, cq, xs, qkv, ffn
    nat.bind(eng.grid, eng._rope_for(eng.grid))
    for kind, slot in (("miss", 0), ("miss", 1), ("hit", 0), ("hit", 1)):
        ours = _engine_lines(eng, kind, slot)
        plan = nat.plan(skip=(kind == "hit"))
        assert len(plan) == len(ours) == (8 if kind == "hit" else 12 + 16 * cfg["num_layers"]), (kind, len(plan), len(ours))
        for i, (a, b) in enumerate(zip(plan, ours)):
            assert a == b, (kind, i, a, b)
        assert "?" not in "".join(plan)
        assert nat.launches(kind == "hit") >= len(plan)
    nat.close()


def test_native_handle_argument_checks():
    """No device work: creation / binding validate their arguments and report through mc_last_error like the rest of the ABI."""
    lib = _lib.lib
    model = wan_ref.WanModel(dim=256, ffn_dim=512, num_heads=2, num_layers=1, text_dim=128, text_len=32).init_synthetic(0)
    weights = mc.WanWeights.from_module(model, torch.device("cpu"))
    nat = native.NativeWanForward(weights)
    assert nat.workspace_bytes((2, 4, 4)) % 1024 == 0 and nat.workspace_bytes((4, 4, 4)) > nat.workspace_bytes((2, 4, 4))
    rope = wan_mod.rope_table((2, 4, 4), 128, "cpu")
    need = ctypes.c_int64(0)
    assert lib.mc_dit_plan(nat.h, 0, None, 0, ctypes.byref(need)) == _lib.MC_ERR_STATE and b"mc_dit_bind" in lib.mc_last_error()
    assert lib.mc_dit_forward(nat.h, 16, 16, 16, 0, 16, 16, None) == _lib.MC_ERR_STATE
    buf = torch.empty(nat.workspace_bytes((2, 4, 4)) + 1024, dtype=torch.uint8)
    ptr = (buf.data_ptr() + 1023) // 1024 * 1024
    assert lib.mc_dit_bind(nat.h, 2, 4, 4, ptr + 8, buf.numel(), rope.data_ptr()) == _lib.MC_ERR_INVALID       # misaligned workspace
    assert lib.mc_dit_bind(nat.h, 2, 4, 4, ptr, 4096, rope.data_ptr()) == _lib.MC_ERR_INVALID                   # too small
    assert b"needed" in lib.mc_last_error()
    assert lib.mc_dit_bind(nat.h, 2, 4, 4, ptr, buf.numel() - 1024, rope.data_ptr()) == _lib.MC_OK
    assert lib.mc_dit_forward(nat.h, None, 16, 16, 0, 16, 16, None) == _lib.MC_ERR_INVALID                      # null latent
    nat.close()
    # unsupported models are refused at creation: 48 output channels (TI2V-5B), i2v
    with pytest.raises(NotImplementedError):
        native.NativeWanForward(types.SimpleNamespace(dims=mc.WAN_CONFIGS["ti2v-5B"]))
    with pytest.raises(NotImplementedError):
        native.NativeWanForward(types.SimpleNamespace(dims=mc.WAN_CONFIGS["i2v-14B"]))
    bad = _lib.DitDims(250, 512, 2, 1, 16, 16, 256, 128, 32, 1e-6)
    wts = _lib.DitWeights()
    wts.blocks = ctypes.cast((_lib.DitBlock * 1)(), ctypes.POINTER(_lib.DitBlock))
    assert not lib.mc_dit_create(ctypes.byref(bad), ctypes.byref(wts)) and b"unsupported dims" in lib.mc_last_error()
    ok = _lib.DitDims(256, 512, 2, 1, 16, 16, 256, 128, 32, 1e-6)
    assert not lib.mc_dit_create(ctypes.byref(ok), ctypes.byref(wts)) and b"null or not 16-byte aligned" in lib.mc_last_error()


def test_nccl_gather_abi_argument_checks():
    """`mc_nccl_*` / `mc_allgather_kv` without a GPU: the entry points exist, validate their arguments and report through
    mc_last_error; a communicator is never created here (that is a collective over GPUs: tests/test_shard_gpu.py, 2 GPUs)."""
    lib = _lib.lib
    assert lib.mc_nccl_unique_id(None) == _lib.MC_ERR_INVALID
    uid = ctypes.create_string_buffer(128)
    rc = lib.mc_nccl_unique_id(uid)
    assert rc in (_lib.MC_OK, _lib.MC_ERR_STATE, _lib.MC_ERR_CUDA)  # OK when a libnccl.so.2 is loadable on this host
    if rc == _lib.MC_OK:
        assert any(uid.raw)
    assert not lib.mc_nccl_init(0, 2, None) and b"null id" in lib.mc_last_error()
    assert not lib.mc_nccl_init(3, 2, uid) and b"rank 3 of 2" in lib.mc_last_error()
    assert lib.mc_allgather_kv(None, 16, None, 16, None, 8, None) == _lib.MC_ERR_INVALID and b"null communicator" in lib.mc_last_error()
    assert lib.mc_nccl_destroy(None) == _lib.MC_OK


def test_native_plan_at_the_benchmarked_shape():
    """The plan of `mc_dit_forward` for BASELINE configs[1] itself (Wan2.1-T2V-1.3B, 832x480x81f: 32 760 tokens, dim 1536, ffn 8960,
    12 heads, 30 layers) — no weights are needed to PLAN, so the weight pointers are stand-in addresses. Checks the launch count, that
    every operand resolves to a known buffer inside its bounds, the GEMM shapes / leading dimensions / epilogues per layer, the
    algorithmic FLOPs the plan adds up to (283.0 TF per forward) and the workspace size."""
    lib = _lib.lib
    D, Fd, H, L, N, TL, TD = 1536, 8960, 12, 30, 21 * 30 * 52, 512, 4096
    dims = _lib.DitDims(D, Fd, H, L, 16, 16, 256, TD, TL, 1e-6)
    addr = iter(range(1 << 44, 1 << 46, 1 << 32))  # distinct, 16-byte aligned, 4 GB apart: never dereferenced by mc_dit_plan
    blocks = (_lib.DitBlock * L)()
    for b in blocks:
        for name in _lib.DIT_BLOCK_FIELDS:
            setattr(b, name, next(addr))
    w = _lib.DitWeights()
    for name in _lib.DIT_TOP_FIELDS:
        setattr(w, name, next(addr))
    w.blocks = ctypes.cast(blocks, ctypes.POINTER(_lib.DitBlock))
    h = lib.mc_dit_create(ctypes.byref(dims), ctypes.byref(w))
    assert h, lib.mc_last_error()
    need = ctypes.c_int64(0)
    _lib.check(lib.mc_dit_workspace_bytes(h, 21, 30, 52, ctypes.byref(need)))
    act = N * (64 * 2 + D * 2 * 4 + D * 4 + 3 * D * 2 + Fd * 2)  # tok, x0 / h / att / cq, xs, qkv, ffn
    assert act < need.value < act + (64 << 20), (need.value, act)      # + text / time buffers, head workspace, split-KV partials
    _lib.check(lib.mc_dit_bind(h, 21, 30, 52, 1 << 40, need.value, next(addr)))
    plans = {}
    for skip in (0, 1):
        n = ctypes.c_int64(0)
        _lib.check(lib.mc_dit_plan(h, skip, None, 0, ctypes.byref(n)))
        buf = ctypes.create_string_buffer(n.value)
        _lib.check(lib.mc_dit_plan(h, skip, buf, n.value, ctypes.byref(n)))
        plans[skip] = buf.value.decode().splitlines()
    lib.mc_dit_destroy(h)
    miss, hit = plans[0], plans[1]
    assert len(miss) == 12 + 16 * L and len(hit) == 8 and hit[:7] == miss[:7] and "?" not in "".join(miss + hit)
    assert hit[7] == f"head x=x0+0:1 r=residual+0 rows={N} cols={D} prep=head_prep+0 -> out+0" and miss[-1].startswith("head x=xs+0:0 r=null")
    assert miss[-2] == f"residual_sub x_out=xs+0 x_in=x0+0 -> residual+0 n={N * D}"
    flops, gemms = 0.0, []
    for line in miss:
        kv = dict(f.split("=", 1) for f in line.split()[1:] if "=" in f)
        if line.startswith("gemm "):
            M, Nn, K = int(kv["M"]), int(kv["N"]), int(kv["K"])
            assert int(kv["lda"]) >= K and int(kv["ldb"]) >= K and int(kv["ldo"]) >= Nn
            flops += 2.0 * M * Nn * K
            gemms.append((M, Nn, K, int(kv["epi"]), kv["out"].split("+")[0], kv["gate"].split("+")[0]))
        elif line.startswith("attention "):
            flops += 4.0 * int(kv["Lq"]) * int(kv["Lk"]) * D
            assert kv["q"].split("+")[0] in ("qkv", "cq") and int(kv["heads"]) == H
    per_layer = gemms[3:3 + 7]
    assert gemms[:3] == [(N, D, 64, 0, "x0", "null"), (TL, D, TD, 1, "ctx_h", "null"), (TL, D, D, 0, "ctx", "null")]
    assert per_layer == [(N, 3 * D, D, 0, "qkv", "null"), (N, D, D, 2, "xs", "em"), (N, D, D, 0, "cq", "null"), (TL, 2 * D, D, 0, "ckv", "null"),
                         (N, D, D, 2, "xs", "null"), (N, Fd, D, 1, "ffn", "null"), (N, D, Fd, 2, "xs", "em")]
    assert all(gemms[3 + 7 * i:3 + 7 * (i + 1)] == per_layer for i in range(L))
    assert abs(flops / 283.0e12 - 1.0) < 0.01, flops  # SURVEY §8d's figure for this forward
