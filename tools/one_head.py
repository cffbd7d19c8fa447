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

# the caller step on a cache-hit step: two hit heads + mc_cfg_step  vs  hit head + hit head with the step in its epilogue
F, Hp, Wp = grid
lat = torch.randn(16, F, 2 * Hp, 2 * Wp, device="cuda", generator=g)
prep = ops.head_prepare(hm, e, wt, b)
def unfused():
    c = ops.head_unpatchify(sets[0][0], hm, e, wt, b, grid, residual=sets[0][1], prep=prep)
    u = ops.head_unpatchify(sets[1][0], hm, e, wt, b, grid, residual=sets[1][1], prep=prep)
    ops.cfg_step(c, u, 5.0, lat, -0.02, out=lat)
def fused():
    c = ops.head_unpatchify(sets[0][0], hm, e, wt, b, grid, residual=sets[0][1], prep=prep)
    ops.head_unpatchify(sets[1][0], hm, e, wt, b, grid, residual=sets[1][1], prep=prep, step=(c, lat, 5.0, 1.0, -0.02), out=lat)
for rnd in range(2):
    for name, fn in (("two heads + cfg_step", unfused), ("head + head-with-step", fused)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        print(f"hit step, {name}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")
