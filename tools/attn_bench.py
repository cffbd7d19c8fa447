"""Attention kernel A/B timing at the benchmarked shape (1xB200): kernel choice, exponential-emulation fraction, split counts,
against torch SDPA on the same tensors. CUDA events, back-to-back launches after warm-up; interleaved rounds because the box
settles under its power cap after the first seconds (only numbers of the same round compare).
    python tools/attn_bench.py [Lq] [rounds] [Lk]        (Lk = 512: the text cross-attention shape)
Writes gpurun_out/attn_bench.json."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from magcache_b200 import ops  # noqa: E402

dev = torch.device("cuda")
N, D, H = 32760, 1536, 12
Lq = int(sys.argv[1]) if len(sys.argv) > 1 else N
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 3
Lk = int(sys.argv[3]) if len(sys.argv) > 3 else N
g = torch.Generator(device=dev).manual_seed(0)
q = torch.randn(Lq, D, device=dev, generator=g).bfloat16()
kv = torch.randn(Lk, 2 * D, device=dev, generator=g).bfloat16()
k, v = kv[:, :D], kv[:, D:]
out = torch.empty(Lq, D, device=dev, dtype=torch.bfloat16)
flops = 4.0 * Lq * Lk * D


def timeit(fn, iters=6 if Lk >= 4096 else 60, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def ours(env):
    def run():
        for kk, vv in env.items():
            os.environ[kk] = vv
        ops.attention(q, k, v, H, out=out)
        for kk in env:
            os.environ.pop(kk, None)
    return run


qh = q.view(1, Lq, H, 128).transpose(1, 2)
kh = k.contiguous().view(1, Lk, H, 128).transpose(1, 2)
vh = v.contiguous().view(1, Lk, H, 128).transpose(1, 2)
cases = {"sdpa_torch": lambda: torch.nn.functional.scaled_dot_product_attention(qh, kh, vh)}
for emu in (0, 2, 3, 4):
    cases[f"long_emu{emu}"] = ours({"MC_ATTN_EMU": str(emu), "MC_ATTN_KERNEL": "2"})
cases["short_kernel"] = ours({"MC_ATTN_KERNEL": "1"})
if Lq < N and Lk == N:
    for sp in (1, 2, 3, 4):
        cases[f"long_splits{sp}"] = ours({"MC_ATTN_SPLITS": str(sp), "MC_ATTN_KERNEL": "2"})
res = {name: [] for name in cases}
order = list(cases)
for r in range(rounds):
    for name in (order if r % 2 == 0 else order[::-1]):
        try:
            ms = timeit(cases[name])
        except Exception as ex:  # noqa: BLE001
            print(name, "failed:", ex)
            ms = float("nan")
        res[name].append(ms)
        print(f"round {r} {name:16s} {ms:8.3f} ms  {flops / ms / 1e9:8.1f} TFLOP/s", flush=True)
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/attn_bench.json", "w") as f:
    json.dump({"Lq": Lq, "Lk": Lk, "heads": H, "flops": flops, "ms": res}, f, indent=1)
