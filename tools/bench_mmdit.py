"""Steps/s of the FLUX.1-dev (BASELINE configs[0] shape: 1024x1024, 28 steps, E024K5R01) and HunyuanVideo (configs[3]: 720p x 129
frames, 50 steps, E024K6R02) forwards on the MMDiT engine with synthetic device-side weights, cached vs non-cached, CUDA events.
Not the contract bench (bench.py measures the north-star Wan2.1 workload); written for the first full-size runs of these engines.

usage (GPU): python tools/bench_mmdit.py flux|hunyuan [--steps N] [--no-cache] [--tokens-scale F]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import magcache_b200 as mc  # noqa: E402
from magcache_b200 import mmdit, ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("family", choices=["flux", "hunyuan"])
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--no-cache", action="store_true")
    ap.add_argument("--frames", type=int, default=33, help="hunyuan: latent frames (33 = 129 video frames)")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    if args.family == "flux":
        steps = args.steps or 28
        model = mmdit.MMDiTHandle(mmdit.FluxEngine(mmdit.random_flux_weights(dev)))
        mc.init_magcache_flux(model, steps, thresh=1e-9 if args.no_cache else 0.24, K=5, retention_ratio=0.1)
        n_img, n_txt = 4096, 512
        hs = torch.randn(1, n_img, 64, device=dev, generator=g).bfloat16()
        enc = torch.randn(1, n_txt, 4096, device=dev, generator=g).bfloat16()
        pooled = torch.randn(1, 768, device=dev, generator=g).bfloat16()
        img_ids = torch.zeros(n_img, 3, device=dev)
        img_ids[:, 1], img_ids[:, 2] = torch.arange(n_img, device=dev) // 64, torch.arange(n_img, device=dev) % 64
        txt_ids = torch.zeros(n_txt, 3, device=dev)
        gd = torch.tensor([3.5], device=dev)

        def step(i):
            return model(hs, enc, pooled, torch.tensor([1.0 - i / steps], device=dev), img_ids, txt_ids, gd, return_dict=False)[0]
    else:
        steps = args.steps or 50
        model = mmdit.MMDiTHandle(mmdit.HunyuanEngine(mmdit.random_hunyuan_weights(dev)))
        mc.init_magcache_hunyuan(model, steps, thresh=1e-9 if args.no_cache else 0.24, K=6, retention_ratio=0.2)
        grid = (args.frames, 45, 80)  # 720 x 1280 -> 90 x 160 latent -> 45 x 80 patches
        n_img = grid[0] * grid[1] * grid[2]
        x = torch.randn(1, 16, grid[0], 2 * grid[1], 2 * grid[2], device=dev, generator=g).bfloat16()
        txt = torch.randn(1, 256, 4096, device=dev, generator=g).bfloat16()
        mask = torch.zeros(1, 256, dtype=torch.long, device=dev)
        mask[0, :48] = 1
        pooled = torch.randn(1, 768, device=dev, generator=g).bfloat16()
        ang = torch.rand(n_img, 64, device=dev, generator=g) * 6.28
        cos, sin = ang.cos().repeat_interleave(2, dim=1), ang.sin().repeat_interleave(2, dim=1)
        gd = torch.tensor([6000.0], device=dev)

        def step(i):
            return model(x, torch.tensor([1000.0 * (1 - i / steps)], device=dev), txt, mask, pooled, cos, sin, gd, return_dict=False)

    for i in range(3):  # warm-up inside the retention window, then restart the schedule
        step(i)
    mc.reset_magcache(model)
    n0 = ops.LAUNCHES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(json.dumps({"family": args.family, "steps": steps, "cached": not args.no_cache, "steps_per_s": steps / (ms / 1e3), "sec_per_sample": ms / 1e3,
                      "image_tokens": n_img, "gpu_launches": ops.LAUNCHES - n0}))


if __name__ == "__main__":
    main()
