"""HBM-bound row kernels at the benchmarked shape (32760 tokens x 1536): achieved GB/s against their algorithmic bytes.
python tools/rowwise_bench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from magcache_b200 import ops  # noqa: E402

N, D, H = 32760, 1536, 12
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
x32 = torch.randn(N, D, device=dev, generator=g)
em = torch.randn(6, D, device=dev, generator=g) * 0.1
out16 = torch.empty(N, D, device=dev, dtype=torch.bfloat16)
qkv = torch.randn(N, 3 * D, device=dev, generator=g).bfloat16()
cq = torch.randn(N, D, device=dev, generator=g).bfloat16()
w2 = torch.ones(2, D, device=dev)
w1 = torch.ones(D, device=dev)
bias = torch.zeros(D, device=dev)
cs = torch.randn(N, 128, device=dev, generator=g)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()  # L2 flush between timed launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters


cases = {
    "ln_modulate fp32->bf16": (lambda: ops.ln_modulate(x32, em, 1, 0, out=out16), N * D * (4 + 2)),
    "ln_affine fp32->bf16": (lambda: ops.ln_affine(x32, w1, bias, out=out16), N * D * (4 + 2)),
    "rmsnorm_rope q|k in place": (lambda: ops.rmsnorm_rope_segs_(qkv[:, :2 * D], w2, 2, cos_sin=cs), N * 2 * D * 4 + N * 128 * 4),
    "rmsnorm (no rope) in place": (lambda: ops.rmsnorm_rope_(cq, w1), N * D * 4),
    "cache_hit_add bf16": (lambda: ops.cache_hit_add(out16, cq), N * D * 6),
}
for name, (fn, nbytes) in cases.items():
    ms = timeit(fn)
    print(f"{name:28s} {ms * 1e3:8.1f} us   {nbytes / ms / 1e6:8.1f} GB/s")
