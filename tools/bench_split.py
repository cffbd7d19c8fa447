"""Split-KV attention: time and check mc_attn_fwd at the per-rank shapes of token-sharded runs for forced split counts.
usage: python tools/bench_split.py   (GPU)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from magcache_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
heads, Lk = 12, 32760
W = heads * 128
torch.manual_seed(0)
k = torch.randn(Lk, W, device=dev).bfloat16()
vt = torch.randn(W, Lk, device=dev).bfloat16()
for Lq in (4095, 8190, 16380, 32760):
    q = torch.randn(Lq, W, device=dev).bfloat16()
    base = None
    for sp in (1, 2, 3, 4, 6, 0):
        os.environ["MC_ATTN_SPLITS"] = str(sp)
        out = ops.attention(q, k, vt, heads)
        for _ in range(3):
            ops.attention(q, k, vt, heads, out=out) if "out" in ops.attention.__code__.co_varnames else ops.attention(q, k, vt, heads)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        torch.cuda.synchronize()
        ev[0].record()
        n = 10
        for _ in range(n):
            o2 = ops.attention(q, k, vt, heads)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / n
        tf = 4.0 * Lq * Lk * W / ms / 1e9
        if base is None:
            base = out.float()
        d = (out.float() - base).abs().max().item()
        same = torch.equal(o2, out)
        print(f"Lq={Lq:6d} splits={sp} {ms:8.4f} ms {tf:7.1f} TF/s  max|d vs splits=1|={d:.3e} reproducible={same}", flush=True)
