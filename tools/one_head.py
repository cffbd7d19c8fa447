"""A few full-shape head launches (hit and stream form) for ncu captures / quick timing: python tools/one_head.py"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from magcache_b200 import ops  # noqa: E402

N, D = 32760, 1536
g = torch.Generator(device="cuda").manual_seed(0)
sets = [(torch.randn(N, D, device="cuda", generator=g).bfloat16(), torch.randn(N, D, device="cuda", generator=g) * 0.3, torch.randn(N, D, device="cuda", generator=g))
        for _ in range(3)]
hm = torch.randn(2, D, device="cuda", generator=g) / math.sqrt(D)
e = torch.randn(1, D, device="cuda", generator=g) * 0.2
wt = (torch.randn(D, 64, device="cuda", generator=g) * 0.03).contiguous()
b = torch.randn(64, device="cuda", generator=g) * 0.1
grid = (21, 30, 52)
for name, fn, nbytes in (("hit", lambda s: ops.head_unpatchify(s[0], hm, e, wt, b, grid, residual=s[1]), N * D * 6 + N * 256),
                         ("stream", lambda s: ops.head_unpatchify(s[2], hm, e, wt, b, grid), N * D * 4 + N * 256)):
    for i in range(3):
        fn(sets[i])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(12):
        fn(sets[i % 3])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 12
    print(f"head {name}: {ms * 1e3:.1f} us per call (prep + main), {nbytes / ms / 1e6:.0f} GB/s", flush=True)
