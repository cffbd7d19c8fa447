"""CPU tests of the multi-GPU (token-shard) host logic with world_size 2 over gloo: partitioning, K/V row all-gather,
partial-output reduction, statistics reduction, and the RoPE table's global token indexing — checked against the oracle."""
import math
import os
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from magcache_b200.shard import TokenShard, allreduce_stats, gather_rows, sum_partial_outputs
from magcache_b200.wan import rope_table
from oracle import wan_ref

GRID = (2, 4, 6)  # 48 tokens
N = GRID[0] * GRID[1] * GRID[2]


def test_token_shard_partition():
    s0, s1 = TokenShard(0, 2, 32760), TokenShard(7, 8, 32760)
    assert (s0.start, s0.stop, s0.n_local) == (0, 16380, 16380)
    assert (s1.start, s1.stop, s1.n_local) == (7 * 4095, 32760, 4095)
    t = torch.arange(32760 * 2).view(32760, 2)
    assert torch.equal(torch.cat([TokenShard(r, 8, 32760).rows(t) for r in range(8)]), t)
    # token counts that do not divide by the world size: the reference's pad rule (videosys/core/comm.py:373-381) — ceil(N / P) row
    # slots per rank, the trailing slots of the last rank are pad
    sh = [TokenShard(r, 8, 32761) for r in range(8)]
    assert all(s.n_slots == 4096 and s.pad == 7 and s.n_padded == 32768 for s in sh)
    assert [s.n_local for s in sh] == [4096] * 7 + [4096 - 7]
    t = torch.arange(32761 * 2).view(32761, 2)
    assert torch.equal(torch.cat([s.rows(t) for s in sh]), t)
    with pytest.raises(ValueError):  # a rank without a single token
        TokenShard(7, 8, 9)


def test_rope_table_matches_oracle_rope_apply():
    """The product's fp32 cos/sin table (global token index -> (f, h, w)) reproduces the oracle's complex128 rope_apply."""
    tab = rope_table(GRID, 128, "cpu")
    assert tab.shape == (N, 128)
    m = wan_ref.WanModel(**wan_ref.CONFIGS["tiny"], text_dim=64, text_len=8)
    x = torch.randn(1, N, 2, 128)
    ref = wan_ref.rope_apply(x, torch.tensor([GRID]), m.freqs)[0]
    cs = tab.view(N, 1, 64, 2).double()
    xc = x[0].double().view(N, 2, 64, 2)
    out = torch.stack([xc[..., 0] * cs[..., 0] - xc[..., 1] * cs[..., 1], xc[..., 0] * cs[..., 1] + xc[..., 1] * cs[..., 0]], -1).view(N, 2, 128)
    assert torch.allclose(out.float(), ref, rtol=1e-5, atol=1e-5)
    # a shard's slice of the table is exactly the rows of its global token range
    sh = TokenShard(1, 2, N)
    assert torch.equal(sh.rows(tab), tab[N // 2:])


def _worker(rank, world, initfile, results):
    dist.init_process_group("gloo", init_method=f"file://{initfile}", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        model = wan_ref.WanModel(**wan_ref.CONFIGS["tiny"], text_dim=64, text_len=8).init_synthetic(3)
        blk = model.blocks[0].self_attn
        x = torch.randn(1, N, 256)
        grid = torch.tensor([GRID])
        with torch.no_grad():
            full = blk(x, torch.tensor([N]), grid, model.freqs)  # unsharded oracle self-attention
            sh = TokenShard(rank, world, N)
            xl = sh.rows(x[0])[None]
            # local projections + norms, RoPE with the GLOBAL positions of this rank's tokens
            q = blk.norm_q(wan_ref.linear_autocast(xl, blk.q))[0].float()
            k = blk.norm_k(wan_ref.linear_autocast(xl, blk.k))[0].float()
            v = wan_ref.linear_autocast(xl, blk.v)[0]
            tab = sh.rows(rope_table(GRID, 128, "cpu")).view(sh.n_local, 1, 64, 2)

            def rope(t):
                tc = t.view(sh.n_local, 2, 64, 2)
                return torch.stack([tc[..., 0] * tab[..., 0] - tc[..., 1] * tab[..., 1], tc[..., 0] * tab[..., 1] + tc[..., 1] * tab[..., 0]],
                                   -1).view(sh.n_local, 256)

            q, k = rope(q).bfloat16(), rope(k).bfloat16()
            k_all, v_all = torch.empty(N, 256, dtype=torch.bfloat16), torch.empty(N, 256, dtype=torch.bfloat16)
            w1 = gather_rows(k.contiguous(), k_all, async_op=True)
            w2 = gather_rows(v.contiguous(), v_all, async_op=True)
            w1.wait(), w2.wait()
            att = wan_ref.attention_ref(q.view(1, -1, 2, 128), k_all.view(1, N, 2, 128), v_all.view(1, N, 2, 128))
            out_l = wan_ref.linear_autocast(att.flatten(2), blk.o)[0]
            err = float((out_l.float() - sh.rows(full[0]).float()).abs().max())
            # partial head outputs: every rank fills only its tokens' positions of a zeroed tensor
            tokens = torch.arange(N * 64, dtype=torch.float32).view(N, 64)
            ref_out = model.unpatchify([tokens], grid)[0]
            part = torch.zeros_like(ref_out)
            mask_tokens = torch.zeros(N, 64)
            mask_tokens[sh.start:sh.stop] = tokens[sh.start:sh.stop]
            part += model.unpatchify([mask_tokens], grid)[0]
            summed = sum_partial_outputs(part)
            stats = allreduce_stats(torch.tensor([1.0 + rank, 2.0, 3.0, float(sh.n_local)], dtype=torch.float64))
            results[rank] = (err, bool(torch.equal(summed, ref_out)), stats.tolist())
    finally:
        dist.destroy_process_group()


def test_sharded_self_attention_gather_and_reductions_world2():
    with tempfile.TemporaryDirectory() as d:
        mgr = mp.get_context("spawn").Manager()  # fork() from a multi-threaded pytest process can deadlock
        results = mgr.dict()
        mp.spawn(_worker, args=(2, os.path.join(d, "init"), results), nprocs=2, join=True)
        assert set(results.keys()) == {0, 1}
        for r in (0, 1):
            err, same, stats = results[r]
            assert err < 3e-2, err      # bf16 pipeline; q/k RoPE in fp32 vs the oracle's fp64
            assert same                  # sum of partial outputs == full unpatchify, exactly
            assert stats == [3.0, 4.0, 6.0, float(N)]


def _engine_worker(rank, world, initfile, results, kind="t2v", grid=GRID):
    """The REAL engine code path of a token-sharded forward (WanEngine with shard_world=2: local projections, async K/V row
    all-gathers, V transpose, attention over all keys, partial-output all-reduce in the head) with the kernels emulated on CPU
    (tests/emu_ops.py) and gloo as the collective backend."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import emu_ops
    import magcache_b200 as mc
    from magcache_b200 import patch as patch_mod
    from magcache_b200 import wan as wan_mod
    dist.init_process_group("gloo", init_method=f"file://{initfile}", rank=rank, world_size=world)
    try:
        wan_mod.ops = emu_ops
        patch_mod.ops = emu_ops
        torch.Tensor.is_cuda = property(lambda self: True)
        os.environ["MC_GRAPHS"] = "0"
        extra_model = {"i2v": dict(in_dim=36, model_type="i2v", clip_dim=64), "vace": dict(model_type="vace", vace_in_dim=24),
                       "ti2v": dict(in_dim=48, out_dim=48)}.get(kind, {})
        model = wan_ref.WanModel(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=64, text_len=8, **extra_model).init_synthetic(3)
        g = torch.Generator().manual_seed(7)
        shape = [grid[0], 2 * grid[1], 2 * grid[2]]
        n_tok = grid[0] * grid[1] * grid[2]
        lat, ctx = torch.randn(48 if kind == "ti2v" else 16, *shape, generator=g), torch.randn(5, 64, generator=g)
        t = torch.tensor([500.0])
        if kind == "ti2v":  # per-token timesteps (Wan2.2 TI2V-5B): a clean range that ends inside rank 1's rows
            t = torch.full((1, n_tok), 500.0)
            t[0, :n_tok // 2 + 6] = 0.0
        extra = {}
        if kind == "i2v":
            extra = dict(clip_fea=torch.randn(1, 257, 64, generator=g), y=[torch.randn(20, *shape, generator=g)])
        if kind == "vace":
            extra = dict(vace_context=[torch.randn(24, *shape, generator=g)])
        table = [1.0] * 8
        outs = {}
        for name, kw in (("single", {}), ("sharded", dict(shard_world=world, shard_rank=rank))):
            m = type("M", (), {})()
            m.model_type = kind
            object.__setattr__(m, "_mc_engine", mc.WanEngine(mc.WanWeights.from_module(model, torch.device("cpu")), **kw))
            mc.init_magcache(m, 4, thresh=10.0, K=3, retention_ratio=0.25, mag_ratios=table)
            seq = []
            with torch.no_grad():
                for i in range(4):  # miss, miss, hit, hit
                    if kind == "vace":
                        seq.append(m.forward([lat], t, extra["vace_context"], [ctx], n_tok)[0].clone())
                    else:
                        seq.append(m.forward([lat], t, [ctx], n_tok, **extra)[0].clone())
            outs[name] = (seq, m._mc_engine, m.residual_cache)
        eng = outs["sharded"][1]
        sh = eng.shard
        errs = [float((a - b).abs().max() / b.abs().max()) for a, b in zip(outs["sharded"][0], outs["single"][0])]
        r_full = outs["single"][2][0][0]
        r_loc = outs["sharded"][2][0][0]
        cache_err = float((r_loc - r_full[sh.start:sh.stop]).abs().max() / r_full.abs().max())
        if kind == "ti2v":
            h = n_tok // 2
            assert eng.t_values == 2 and eng.runs == ([(0, h, 0)] if rank == 0 else [(0, 6, 0), (6, h, 1)]), eng.runs
            assert outs["single"][1].runs == [(0, h + 6, 0), (h + 6, n_tok, 1)] and seq[0].shape[0] == 48
        results[rank] = (errs, cache_err, tuple(r_loc.shape), (sh.start, sh.stop))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["t2v", "i2v", "vace", "ti2v"])
def test_sharded_engine_equals_single_engine_world2(kind):
    with tempfile.TemporaryDirectory() as d:
        mgr = mp.get_context("spawn").Manager()
        results = mgr.dict()
        mp.spawn(_engine_worker, args=(2, os.path.join(d, "init"), results, kind), nprocs=2, join=True)
        assert set(results.keys()) == {0, 1}
        for r in (0, 1):
            errs, cache_err, shape, rng = results[r]
            assert len(errs) == 4 and max(errs) < 2e-3, errs       # same arithmetic on row subsets (CPU matmul blocking differs)
            assert cache_err < 2e-3 and shape == (N // 2, 256) and rng == (r * N // 2, (r + 1) * N // 2)


def test_sharded_engine_with_padded_token_count_world2():
    """45 tokens over 2 ranks: 23 row slots each, rank 1 owns 22 tokens + 1 pad slot (the reference's pad rule,
    videosys/core/comm.py:373-381); outputs and the sharded residual cache equal the single engine's."""
    grid = (1, 5, 9)
    with tempfile.TemporaryDirectory() as d:
        mgr = mp.get_context("spawn").Manager()
        results = mgr.dict()
        mp.spawn(_engine_worker, args=(2, os.path.join(d, "init"), results, "t2v", grid), nprocs=2, join=True)
        for r, (lo, hi) in enumerate(((0, 23), (23, 45))):
            errs, cache_err, shape, rng = results[r]
            assert len(errs) == 4 and max(errs) < 2e-3, errs
            assert cache_err < 2e-3 and shape == (hi - lo, 256) and rng == (lo, hi)
