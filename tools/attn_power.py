"""Self-attention at the benchmarked shape, ours vs torch SDPA, each looped for a few seconds while nvidia-smi samples SM clock and
board power: tells a cycle-count gap from a power-efficiency gap on a power-capped board.   python tools/attn_power.py [seconds]"""
import os
import subprocess
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from magcache_b200 import ops  # noqa: E402

N, D, H = 32760, 1536, 12
secs = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(N, 3 * D, device="cuda", generator=g).bfloat16()
q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
out = torch.empty(N, D, device="cuda", dtype=torch.bfloat16)
qh, kh, vh = (t.contiguous().view(1, N, H, 128).transpose(1, 2) for t in (q, k, v))
flops = 4.0 * N * N * D


def sample(stop, rows):
    while not stop.is_set():
        r = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.active", "--format=csv,noheader,nounits", "-i", "0"],
                           capture_output=True, text=True).stdout.strip().split(",")
        try:
            rows.append((float(r[0]), float(r[1])))
        except (ValueError, IndexError):
            pass
        time.sleep(0.1)


def run(name, fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    stop, rows = threading.Event(), []
    th = threading.Thread(target=sample, args=(stop, rows))
    th.start()
    t0 = time.time()
    n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < secs:
        for _ in range(10):
            fn()
        n += 10
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    stop.set()
    th.join()
    ms = e0.elapsed_time(e1) / n
    rows = rows[len(rows) // 3:]  # settled part
    clk = sorted(r[0] for r in rows)[len(rows) // 2] if rows else float("nan")
    pw = sorted(r[1] for r in rows)[len(rows) // 2] if rows else float("nan")
    print(f"{name:24s} {ms:7.3f} ms  {flops / ms / 1e9:7.1f} TF/s   SM clock {clk:6.0f} MHz  power {pw:6.0f} W   cycles/launch {ms * clk * 1e3 / 1e6:8.2f} M   J/launch {ms * pw / 1e3:6.2f}",
          flush=True)


for emu in ("0", "2", "3"):
    os.environ["MC_ATTN_EMU"] = emu
    run(f"ours (emu {emu}/8)", lambda: ops.attention(q, k, v, H, out=out))
os.environ.pop("MC_ATTN_EMU")
run("torch SDPA (cuDNN)", lambda: torch.nn.functional.scaled_dot_product_attention(qh, kh, vh))
run("ours (default)", lambda: ops.attention(q, k, v, H, out=out))
