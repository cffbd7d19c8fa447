"""One full-shape self-attention launch (for ncu captures): python tools/one_attn.py [Lq]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from magcache_b200 import ops  # noqa: E402

N, D, H = 32760, 1536, 12
Lq = int(sys.argv[1]) if len(sys.argv) > 1 else N
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(N, 3 * D, device="cuda", generator=g).bfloat16()
out = torch.empty(Lq, D, device="cuda", dtype=torch.bfloat16)
for _ in range(2):
    ops.attention(qkv[:Lq, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], H, out=out)
torch.cuda.synchronize()
print("done", float(out.float().abs().mean()))
