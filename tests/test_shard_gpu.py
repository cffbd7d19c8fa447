"""Multi-GPU parity: the token-sharded forward (world 2) against the single-GPU forward of the same engine, for both exchanges —
the peer-to-peer one (copy-engine pushes + flags consumed inside the attention kernel, head rows stored into the peer's output)
and the collective one (NCCL all-gather / all-reduce, MC_SHARD_P2P=0) — over a miss, miss, hit, hit sequence, with CUDA-graph
replay on and off, for a token count that divides by the world size and one that needs the pad rule.
Every kernel's per-row arithmetic is independent of how the rows are partitioned; only the attention's work decomposition differs
(256-row CTAs, rotated key order), so outputs agree to the rounding of P, not bit for bit.
Needs >= 2 GPUs (run with `gpurun --gpus 2`); skipped on a 1-GPU box."""
import os
import tempfile

import pytest
import torch

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def _worker(rank, world, initfile, results, latent_shape):
    import torch.distributed as dist

    import magcache_b200 as mc
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", init_method=f"file://{initfile}", rank=rank, world_size=world, device_id=dev)
    try:
        dims = mc.WanDims(dim=256, ffn_dim=512, num_heads=2, num_layers=3, text_dim=128, text_len=32)
        g = torch.Generator().manual_seed(0)
        lat = torch.randn(*latent_shape, generator=g).to(dev)
        n_tok = latent_shape[1] * (latent_shape[2] // 2) * (latent_shape[3] // 2)
        ctx = torch.randn(20, 128, generator=g).bfloat16().to(dev)
        t = torch.tensor([640.0], device=dev)
        table = [1.0, 1.0] + [0.999] * 18
        outs = {}
        for mode in ("single", "p2p", "p2p_graphs", "collective"):
            os.environ["MC_SHARD_P2P"] = "0" if mode == "collective" else "1"
            os.environ["MC_GRAPHS"] = "1" if mode == "p2p_graphs" else "0"
            w = mc.WanWeights.random(dims, dev, seed=5)
            m = mc.WanModelHandle(w) if mode == "single" else mc.WanModelHandle(w, shard_world=world, shard_rank=rank)
            mc.init_magcache(m, 10, thresh=0.12, K=2, retention_ratio=0.1, mag_ratios=table)
            seq = []
            for rep in range(3 if mode == "p2p_graphs" else 1):  # graphs: eager, capture, replay
                mc.reset_magcache(m)
                seq = [m([lat], t=t, context=[ctx], seq_len=n_tok)[0].clone() for _ in range(4)]  # miss, miss, hit, hit
            outs[mode] = (seq, m.residual_cache[0].clone(), m._mc_engine)
            torch.cuda.synchronize()
            dist.barrier()
        ref_seq, ref_cache, _ = outs["single"]
        res = {}
        for mode in ("p2p", "p2p_graphs", "collective"):
            seq, cache, eng = outs[mode]
            sh = eng.shard
            errs = [rel_l2(a, b) for a, b in zip(seq, ref_seq)]
            c_err = rel_l2(cache.view(sh.n_local, 256), ref_cache.view(n_tok, 256)[sh.start:sh.stop])
            res[mode] = (max(errs), c_err, type(eng.xch).__name__)
        results[rank] = (res, float(ref_seq[0].abs().mean()))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("latent_shape", [(16, 5, 32, 48), (16, 3, 18, 30)])  # 1920 tokens (960 per rank); 405 tokens (203 + 202: pad rule)
def test_sharded_forward_matches_single_gpu(latent_shape):
    import torch.multiprocessing as mp
    with tempfile.TemporaryDirectory() as d:
        mgr = mp.Manager()
        results = mgr.dict()
        mp.spawn(_worker, args=(2, os.path.join(d, "init"), results, latent_shape), nprocs=2, join=True)
        assert set(results.keys()) == {0, 1}
        for r in (0, 1):
            res, mag = results[r]
            assert mag > 0
            assert res["p2p"][2] == "P2PExchange" and res["collective"][2] == "CollectiveExchange"
            for mode, (err, c_err, _) in res.items():
                assert err < 3e-3, (mode, "outputs", err)
                assert c_err < 3e-3, (mode, "residual cache slice", c_err)
