"""Wan2.2 TI2V-5B (MagCache4Wan2.2/magcache_generate.py:209-336; dim 3072, 24 heads, ffn 14336, 30 layers, 48 latent channels) at its
full 1280x704x121f shape (latent 48 x 31 x 44 x 80 -> 27 280 tokens) on the Wan engine with synthetic device-side weights: per-forward
CUDA-event times of cache misses and hits for the image-to-video form (first-frame tokens at t = 0: per-token timesteps, two row
ranges) and the text-to-video form (one timestep), preset E006K2R02 walked from step 8. Not the contract bench (bench.py measures
the north-star Wan2.1 workload).

usage (GPU): python tools/bench_ti2v.py [--steps 8] [--frames 31]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import magcache_b200 as mc  # noqa: E402
from magcache_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--frames", type=int, default=31)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    dims = mc.WAN_CONFIGS["ti2v-5B"]
    grid = (args.frames, 22, 40)
    n_tok = grid[0] * grid[1] * grid[2]
    weights = mc.WanWeights.random(dims, dev, seed=0)
    g = torch.Generator(device=dev).manual_seed(0)
    lat = torch.randn(48, grid[0], 2 * grid[1], 2 * grid[2], device=dev, generator=g)
    ctxs = [torch.randn(512, 4096, device=dev, generator=g).bfloat16() for _ in range(2)]
    sample_steps, start = 50, 8
    s = torch.linspace(1.0, 1.0 / sample_steps, sample_steps)
    sig = 5.0 * s / (1 + 4.0 * s)
    attn_flops = 4.0 * n_tok * n_tok * dims.dim
    lin_flops = 2.0 * n_tok * (4 * dims.dim ** 2 + 2 * dims.dim * dims.ffn_dim + 2 * dims.dim ** 2) + 2.0 * 512 * 2 * dims.dim ** 2 \
        + 4.0 * n_tok * 512 * dims.dim
    fwd_flops = dims.num_layers * (attn_flops + lin_flops)
    out = {"workload": f"Wan2.2 TI2V-5B, latent 48x{grid[0]}x{2 * grid[1]}x{2 * grid[2]}", "tokens": n_tok, "preset": "E006K2R02, steps 8.. of 50",
           "forward_tflop": fwd_flops / 1e12}
    for form in ("i2v_per_token_t", "t2v_one_t"):
        model = mc.WanModelHandle(weights)
        mc.init_magcache_wan22(model, "wan2.2_ti2v_5b_a", sample_steps, thresh=0.06, K=2, retention_ratio=0.2)
        ev, kinds = [], []
        n0 = ops.LAUNCHES
        for rep in ("warm", "timed"):
            type(model).cnt = torch.tensor(2 * start)
            type(model).accumulated_err, type(model).accumulated_steps, type(model).accumulated_ratio = [0.0, 0.0], [0, 0], [1.0, 1.0]
            ev.clear(), kinds.clear()
            n0 = ops.LAUNCHES
            for i in range(start, start + (3 if rep == "warm" else args.steps)):
                t = torch.full((1, n_tok), float(1000.0 * sig[i]), device=dev)
                if form == "i2v_per_token_t":
                    t[0, :grid[1] * grid[2]] = 0.0
                for c in ctxs:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    y = model([lat], t=t, context=[c], seq_len=n_tok)[0]
                    e1.record()
                    ev.append((e0, e1))
                    # a miss leaves the slot's skip count at 0 (inside the retention window it never moves), a hit raises it
                    kinds.append("hit" if type(model).accumulated_steps[(int(type(model).cnt) - 1) % 2] > 0 else "miss")
            torch.cuda.synchronize()
        ms = [a.elapsed_time(b) for a, b in ev]
        miss = [m for m, k in zip(ms, kinds) if k == "miss"]
        hit = [m for m, k in zip(ms, kinds) if k == "hit"]
        eng = model._mc_engine
        out[form] = {"forwards": {"miss": len(miss), "hit": len(hit)}, "miss_ms": sum(miss) / max(1, len(miss)),
                     "hit_ms": (sum(hit) / len(hit)) if hit else None, "miss_tflops": fwd_flops / 1e9 / (sum(miss) / max(1, len(miss))),
                     "row_ranges": eng.runs, "head_launches_per_forward": len(eng.head_groups) * (len(eng.runs) if eng.runs else 1),
                     "gpu_launches": ops.LAUNCHES - n0, "finite": bool(torch.isfinite(y).all())}
        del model
    print(json.dumps(out))


if __name__ == "__main__":
    main()
