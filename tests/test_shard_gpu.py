"""Multi-GPU parity: the token-sharded forward (world 2, NCCL) equals the single-GPU forward bit for bit — every kernel's
per-element arithmetic is independent of how the token rows are partitioned, and the partial-output reduction adds zeros.
Needs >= 2 GPUs (run with `gpurun --gpus 2`); skipped on a 1-GPU box."""
import os
import tempfile

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, initfile, results):
    import torch.distributed as dist

    import magcache_b200 as mc
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", init_method=f"file://{initfile}", rank=rank, world_size=world, device_id=dev)
    try:
        dims = mc.WanDims(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=128, text_len=32)
        g = torch.Generator().manual_seed(0)
        lat = torch.randn(16, 2, 16, 16, generator=g).to(dev)   # 128 tokens -> 64 per rank
        ctx = torch.randn(20, 128, generator=g).bfloat16().to(dev)
        t = torch.tensor([640.0], device=dev)
        table = [1.0, 1.0] + [0.999] * 18
        outs = {}
        for mode in ("sharded", "single"):
            w = mc.WanWeights.random(dims, dev, seed=5)
            m = mc.WanModelHandle(w, shard_world=world, shard_rank=rank) if mode == "sharded" else mc.WanModelHandle(w)
            mc.init_magcache(m, 10, thresh=0.12, K=2, retention_ratio=0.1, mag_ratios=table)
            seq = []
            for _ in range(4):  # miss, miss, hit, hit
                seq.append(m([lat], t=t, context=[ctx], seq_len=128)[0].clone())
            outs[mode] = (seq, m.residual_cache[0].clone())
        same = all(torch.equal(a, b) for a, b in zip(outs["sharded"][0], outs["single"][0]))
        n_local = 128 // world
        r_full = outs["single"][1].view(128, 256)[rank * n_local:(rank + 1) * n_local]
        cache_same = torch.equal(outs["sharded"][1].view(n_local, 256), r_full)
        results[rank] = (bool(same), bool(cache_same), float(outs["single"][0][0].abs().mean()))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_forward_equals_single_gpu_bitwise():
    import torch.multiprocessing as mp
    with tempfile.TemporaryDirectory() as d:
        mgr = mp.Manager()
        results = mgr.dict()
        mp.spawn(_worker, args=(2, os.path.join(d, "init"), results), nprocs=2, join=True)
        assert set(results.keys()) == {0, 1}
        for r in (0, 1):
            same, cache_same, mag = results[r]
            assert mag > 0
            assert same, "sharded outputs differ from the single-GPU outputs"
            assert cache_same, "sharded residual cache is not the rank's slice of the single-GPU cache"
