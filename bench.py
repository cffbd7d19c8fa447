#!/usr/bin/env python
"""bench.py — denoising steps/sec of the MagCache hot path on Wan2.1-T2V-1.3B, 832x480x81 frames (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--no-cache]

One "step" = one denoising step = the cond + uncond pair of patched-forward calls the reference's caller makes
(eval/magcache/experiments/Wan2.1_EVAL/wan_magcache.py:296-299) under the E012K4R02 schedule
(thresh 0.12, K 4, retention 0.2: MagCache4Wan2.1/README.md:13), 50-step video => 42 block-stack forwards + 58 cache hits.
The K timed steps walk that schedule from cnt = 0 (warm-up steps are extra, then the controller is reset), so the default
K = 50 times exactly one video. Synthetic latents / text embeddings / seeded random weights of the named architecture
(no checkpoints offline). Prints ONE JSON line (see the task contract): value = steps/sec with inputs resident in HBM,
e2e = the same through the public `model(x, t, context, seq_len)` call with host tensors (H2D + D2H inside the timed region),
roofline = the dominant kernel (self-attention) from CUDA events recorded live around each of its launches,
cpu_baseline = the CPU oracle timed on this box's host cores on a bounded sample.

`--impl reference` times the reference's own path on the host CPU: the reference is pure Python/torch whose model code
(`wan`) is not installable offline, so the arm runs the oracle restatement (oracle/wan_ref.py, "port") with all host threads.
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GRID = (21, 30, 52)          # latent 16 x 21 x 60 x 104  -> 32760 tokens
LATENT = (16, 21, 60, 104)
SAMPLE_STEPS = 50
PRESET = dict(thresh=0.12, K=4, retention_ratio=0.2)
N_TOK = GRID[0] * GRID[1] * GRID[2]
D, FFN, HEADS, LAYERS, TEXT_LEN, TEXT_DIM = 1536, 8960, 12, 30, 512, 4096
ATTN_SELF_FLOPS = 4.0 * N_TOK * N_TOK * D  # QK^T + PV, 2 flop per MAC (6.594 TF per layer)
LAYER_FLOPS = 9.433e12
FWD_FLOPS = 283.0e12
WORKLOAD = "Wan2.1-T2V-1.3B 832x480x81f"
MODEL_KEY, TABLE = "t2v-1.3B", "wan2.1_t2v_1.3b"


def select_workload(name):
    """Default = BASELINE configs[1]. `wan14b` = configs[4]'s model and resolution (Wan2.1-T2V-14B, 1280x720x81f, E024K6R02):
    not what the driver times, kept to show the 14B shapes run at full size (one GPU holds it: 28 GB of bf16 weights)."""
    global GRID, LATENT, PRESET, N_TOK, D, FFN, HEADS, LAYERS, ATTN_SELF_FLOPS, FWD_FLOPS, WORKLOAD, MODEL_KEY, TABLE
    if name == "wan14b":
        GRID, LATENT = (21, 45, 80), (16, 21, 90, 160)
        PRESET = dict(thresh=0.24, K=6, retention_ratio=0.2)
        N_TOK = GRID[0] * GRID[1] * GRID[2]
        D, FFN, HEADS, LAYERS = 5120, 13824, 40, 40
        ATTN_SELF_FLOPS = 4.0 * N_TOK * N_TOK * D
        FWD_FLOPS = 6523.0e12
        WORKLOAD, MODEL_KEY, TABLE = "Wan2.1-T2V-14B 1280x720x81f", "t2v-14B", "wan2.1_t2v_14b"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm_gbs=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
    return dict(hbm_gbs=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index), "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
        busy = [v for v in sm if mx and v > 0.3 * mx] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm (oracle "port" of the reference path) — bounded sample, extrapolated; states exactly what was timed
# ---------------------------------------------------------------------------------------------------------------------
def host_cores():
    """CPU threads this process may really use: affinity mask, clipped by the cgroup CPU quota when one is set."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:  # noqa: BLE001
        n = os.cpu_count() or 1
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:
            quota, period = f.read().split()
        if quota != "max":
            n = max(1, min(n, int(math.ceil(int(quota) / int(period)))))
    except Exception:  # noqa: BLE001
        pass
    return n


def pick_threads(state):
    """torch-CPU scales badly past the physical cores the container really owns: try a few thread counts on one sample
    each and keep the fastest ("all the host threads it can use", not more)."""
    import torch
    n = host_cores()
    best, best_t = n, None
    for c in sorted({min(n, v) for v in (8, 16, 32, 64, n)}):
        torch.set_num_threads(c)
        cpu_time_block(state)
        t, _ = cpu_time_block(state)
        if best_t is None or t < best_t:
            best, best_t = c, t
    torch.set_num_threads(best)
    return best


def cpu_sample_setup(n_frames=21, hp=15, wp=13):
    import torch
    from oracle import wan_ref
    torch.manual_seed(0)
    blk = wan_ref.WanAttentionBlock(D, FFN, HEADS).eval()
    model_bits = dict(blk=blk, freqs=wan_ref.WanModel(dim=D, ffn_dim=16, num_heads=HEADS, num_layers=0).freqs)
    n = n_frames * hp * wp
    x = torch.randn(1, n, D)
    e = torch.randn(1, 6, D) * 0.1
    ctx = torch.randn(1, TEXT_LEN, D).bfloat16()
    grid = torch.tensor([[n_frames, hp, wp]])
    return wan_ref, model_bits, x, e, ctx, grid, n


def cpu_time_block(state):
    """Seconds for ONE WanAttentionBlock forward at the sample size, split into (attention, everything else)."""
    import torch
    wan_ref, mb, x, e, ctx, grid, n = state
    blk = mb["blk"]
    with torch.no_grad():
        t0 = time.perf_counter()
        blk(x, e, torch.tensor([n]), grid, mb["freqs"], ctx, None)
        t_block = time.perf_counter() - t0
        q = torch.randn(1, n, HEADS, 128).bfloat16()
        t0 = time.perf_counter()
        wan_ref.attention_ref(q, q, q)
        t_attn = time.perf_counter() - t0
    return t_block, t_attn


def cpu_extrapolate(t_block, t_attn, n_sample):
    """Full-shape forward time from the sample: attention scales with N^2, the rest with N (prologue/head << 1 block)."""
    s = N_TOK / n_sample
    t_full_block = (t_block - t_attn) * s + t_attn * s * s
    t_miss = LAYERS * t_full_block
    t_hit = 0.02 * t_full_block  # prologue + add + head: ~3 passes over [N, D] and a 64-wide GEMM (bounded above by 2% of a block)
    sec_video = 42 * t_miss + 58 * t_hit
    return SAMPLE_STEPS / sec_video, sec_video, t_miss


def run_reference_arm(args, rank):
    import torch
    if rank != 0:
        return
    state = cpu_sample_setup()
    n_sample = state[-1]
    cores = pick_threads(state)
    for _ in range(max(1, min(args.warmup, 2))):
        cpu_time_block(state)
    tb, ta = [], []
    t_start = time.perf_counter()
    for _ in range(args.steps):
        b, a = cpu_time_block(state)
        tb.append(b)
        ta.append(a)
        if time.perf_counter() - t_start > 150:  # keep the whole run within a few minutes
            break
    t_block, t_attn = statistics.median(tb), statistics.median(ta)
    value, sec_video, t_miss = cpu_extrapolate(t_block, t_attn, n_sample)
    sample = (f"{len(tb)} x one WanAttentionBlock (of {LAYERS}) at {n_sample} of {N_TOK} tokens on torch-CPU (bf16-autocast emulation, "
              f"{cores} threads): block {t_block:.3f}s of which SDPA {t_attn:.3f}s; extrapolated attention ~N^2, rest ~N; "
              f"forward = {LAYERS} blocks = {t_miss:.1f}s; video = 42 miss + 58 hit forwards (E012K4R02)")
    line = {"metric": "denoising_steps_per_sec", "value": value, "unit": "steps/s", "n_gpus": 0, "steps": len(tb), "warmup": args.warmup,
            "ms_per_step": 1000.0 / value, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "impl": "reference", "sec_per_video": sec_video,
            "config": {"workload": "Wan2.1-T2V-1.3B 832x480x81f, 50 steps, MagCache E012K4R02 (BASELINE configs[1])", "tokens": N_TOK,
                       "note": "reference = pure-Python/torch path; upstream `wan` not installable offline -> oracle restatement on host CPU"},
            "cpu_baseline": {"value": value, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------------
def run_ours(args, rank, world):
    import torch
    import torch.distributed as dist

    import magcache_b200 as mc
    from magcache_b200 import ops

    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    weights = mc.WanWeights.random(mc.WAN_CONFIGS[MODEL_KEY], dev, seed=0)  # same seed on every rank: replicated weights
    # N > 1: ONE video, token axis sharded over the ranks (K/V all-gather per layer), see magcache_b200/shard.py
    model = mc.WanModelHandle(weights, shard_world=world, shard_rank=rank) if world > 1 else mc.WanModelHandle(weights)
    thresh = 1e-9 if args.no_cache else PRESET["thresh"]  # --no-cache: the controller never skips (same code path, same shapes)
    mc.init_magcache(model, SAMPLE_STEPS, thresh=thresh, K=PRESET["K"], retention_ratio=PRESET["retention_ratio"], table=TABLE)

    g = torch.Generator().manual_seed(0)
    lat_h = torch.randn(*LATENT, generator=g).pin_memory()
    ctx_h = torch.randn(TEXT_LEN, TEXT_DIM, generator=torch.Generator().manual_seed(1)).bfloat16().pin_memory()
    ctxn_h = torch.randn(TEXT_LEN, TEXT_DIM, generator=torch.Generator().manual_seed(2)).bfloat16().pin_memory()
    out_h = torch.empty(2, *LATENT).pin_memory()
    lat_d, ctx_d, ctxn_d = lat_h.to(dev), ctx_h.to(dev), ctxn_h.to(dev)
    shift = 5.0
    s = torch.linspace(1.0, 1.0 / SAMPLE_STEPS, SAMPLE_STEPS)
    sig = torch.cat([shift * s / (1 + (shift - 1) * s), torch.zeros(1)])
    t_dev = [(sig[i:i + 1] * 1000.0).to(dev) for i in range(SAMPLE_STEPS)]
    guide = 5.0

    def reset_controller():
        mc.reset_magcache(model)

    def step_resident(i, x):
        t = t_dev[i % SAMPLE_STEPS]
        cond = model([x], t=t, context=[ctx_d], seq_len=N_TOK)[0]
        uncond = model([x], t=t, context=[ctxn_d], seq_len=N_TOK)[0]
        # caller-side code (wan_magcache.py:301-310): CFG combine + an Euler flow step standing in for FlowUniPC, one fused kernel
        return ops.cfg_step(cond, uncond, guide, x, float(sig[(i % SAMPLE_STEPS) + 1] - sig[i % SAMPLE_STEPS]), out=x)

    def step_e2e(i):
        t = t_dev[i % SAMPLE_STEPS]
        x = lat_h.to(dev, non_blocking=True)
        c = ctx_h.to(dev, non_blocking=True)
        cond = model([x], t=t, context=[c], seq_len=N_TOK)[0]
        out_h[0].copy_(cond, non_blocking=True)
        x2 = lat_h.to(dev, non_blocking=True)
        cn = ctxn_h.to(dev, non_blocking=True)
        uncond = model([x2], t=t, context=[cn], seq_len=N_TOK)[0]
        out_h[1].copy_(uncond, non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(run_step, profile):
        reset_controller()
        x = lat_d.clone()
        for i in range(args.warmup):  # warm-up: extra non-cached steps (cnt < retention window), then restart the schedule
            r = run_step(i, x) if run_step is step_resident else run_step(i)
            x = r if r is not None else x
            if model.cnt >= 20:
                reset_controller()
        reset_controller()
        x = lat_d.clone()
        ops.PROFILE = {} if profile else None
        launches0 = ops.LAUNCHES
        sampler = ClockSampler(local)
        barrier()
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            r = run_step(i, x) if run_step is step_resident else run_step(i)
            x = r if r is not None else x
        e1.record()
        barrier()
        clocks = sampler.stop()
        ms = e0.elapsed_time(e1)
        prof, ops.PROFILE = ops.PROFILE, None
        if world > 1:
            tms = torch.tensor([ms], device=dev)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            ms = float(tms.item())
        return ms, ops.LAUNCHES - launches0, clocks, prof, x

    eng = model._mc_engine
    graphs = eng.use_graphs
    ms, launches, clocks, prof, x_final = timed(step_resident, profile=not graphs)
    assert torch.isfinite(x_final).all(), "non-finite latents after the timed steps"
    ms_e2e, _, _, _, _ = timed(step_e2e, profile=False)
    roofline_source = "CUDA events around every launch inside the timed region"
    if graphs:
        # the timed region replays CUDA graphs (no per-kernel events inside a graph): take the per-kernel times from an eager
        # pass of the first steps of the same schedule, right after the timed passes
        eng.use_graphs = False
        keep = args.steps
        args.steps = min(args.steps, 4)
        _, _, _, prof, _ = timed(step_resident, profile=True)
        args.steps = keep
        eng.use_graphs = True
        roofline_source = f"eager pass of the first {min(keep, 4)} steps with CUDA events around every launch (the timed region replays CUDA graphs)"

    # skip schedule actually walked in the timed region
    from magcache_b200.controller import make_ctrl_config, schedule_mask
    cfgm = mc.MagCacheConfig("wan2.1", thresh, PRESET["K"], PRESET["retention_ratio"], SAMPLE_STEPS, table=TABLE)
    mask = schedule_mask(make_ctrl_config(cfgm.num_steps, cfgm.thresh, cfgm.K, cfgm.retention_ratio, cfgm.resolved_ratios(), **cfgm.ctrl_kwargs()), 2 * SAMPLE_STEPS)
    walked = [int(mask[c % (2 * SAMPLE_STEPS)]) for c in range(2 * args.steps)]
    n_hit, n_miss = sum(walked), len(walked) - sum(walked)

    pk = peaks()
    kern = {}
    for tag, evs in (prof or {}).items():
        ts = [a.elapsed_time(b) for a, b in evs]
        kern[tag] = {"launches": len(ts), "ms_avg": sum(ts) / len(ts), "ms_total": sum(ts)}
    roof = None
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r01_attn_traffic.json")
    if world == 1 and MODEL_KEY == "t2v-1.3B" and os.path.exists(tp):  # dram__bytes_read + write of one full-shape launch, from the committed ncu capture
        with open(tp) as f:
            traffic = json.load(f)["traffic_bytes_per_launch"]
    if "attn_self" in kern:
        ach = (ATTN_SELF_FLOPS / world) / (kern["attn_self"]["ms_avg"] * 1e-3) / 1e12  # per GPU: N/world query rows x N keys
        roof = {"kernel": f"attn_fwd_kernel (self-attention, {N_TOK}x{N_TOK}x{HEADS} heads)", "bound": "tensor", "achieved": ach, "peak": pk["tf_sustained"],
                "unit": "TFLOP/s", "frac": ach / pk["tf_sustained"], "traffic": traffic, "peak_source": pk["src"] + " (sustained bf16)",
                "share_of_step": (kern["attn_self"]["ms_total"] / ms) if not graphs else None, "flops_per_launch": ATTN_SELF_FLOPS / world,
                "measured": roofline_source}
    # the HBM-bound cache-hit add, timed alone on rotating buffers (inputs 3 x 503 MB > L2)
    k1 = bench_k1(dev, pk) if MODEL_KEY == "t2v-1.3B" else None

    steps_per_s = args.steps / (ms * 1e-3)  # whole job: all ranks work on the same video
    e2e_v = args.steps / (ms_e2e * 1e-3)
    h2d = 2 * (lat_h.numel() * 4 + ctx_h.numel() * 2)
    d2h = 2 * lat_h.numel() * 4
    line = {"metric": "denoising_steps_per_sec", "value": steps_per_s, "unit": "steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD + ", 50 steps, MagCache " + ("disabled (non-cached loop)" if args.no_cache else
                                    f"E{int(PRESET['thresh'] * 100):03d}K{PRESET['K']}R{int(PRESET['retention_ratio'] * 10):02d}") +
                                    (" (BASELINE configs[1])" if MODEL_KEY == "t2v-1.3B" else " (BASELINE configs[4] model and shape, single GPU)"),
                       "tokens": N_TOK, "dim": D, "layers": LAYERS, "forwards_timed": {"miss": n_miss, "hit": n_hit},
                       "parallelism": "single GPU" if world == 1 else f"token-axis shard over {world} GPUs ({N_TOK // world} tokens each), NCCL all-gather of K and V per layer, replicated weights",
                       "cuda_graphs": bool(graphs),
                       "l2_policy": "per-forward working set (>= 1.3 GB of activations + 2.8 GB weights) exceeds the 126 MB L2; no explicit flush"},
            "sec_per_video": (ms * 1e-3) * SAMPLE_STEPS / args.steps,
            "e2e": {"value": e2e_v, "unit": "steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof, "kernels": kern, "k1_cache_hit_add": k1,
            "model_flops_per_miss_forward": FWD_FLOPS,
            "achieved_tflops_miss_only_whole_job": (n_miss * FWD_FLOPS / 1e12) / (ms * 1e-3) if n_miss else None}
    if rank == 0:
        if world == 1 and not args.skip_cpu and MODEL_KEY == "t2v-1.3B":
            line["cpu_baseline"] = cpu_baseline_leg()
        print(json.dumps(line), flush=True)
    if world > 1:
        shutdown_distributed(model)


def shutdown_distributed(model):
    """Leave cleanly and in bounded time: captured CUDA graphs hold NCCL work, and tearing the process group down under them
    can block forever — drop the graphs first, and never wait more than a few seconds for the communicator to go away."""
    import gc

    import torch
    import torch.distributed as dist
    eng = getattr(model, "_mc_engine", None)
    if eng is not None:
        eng._graphs.clear()
    gc.collect()
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    done = threading.Event()

    def _destroy():
        try:
            dist.barrier()
            dist.destroy_process_group()
        except Exception:  # noqa: BLE001
            pass
        done.set()

    th = threading.Thread(target=_destroy, daemon=True)
    th.start()
    done.wait(timeout=15)
    os._exit(0)


def bench_k1(dev, pk):
    import torch
    from magcache_b200 import ops
    n = N_TOK * D
    sets = [(torch.randn(n, device=dev).bfloat16(), torch.randn(n, device=dev), torch.empty(n, device=dev)) for _ in range(3)]
    for i in range(6):
        ops.cache_hit_add(sets[i % 3][0], sets[i % 3][1], out=sets[i % 3][2])
    torch.cuda.synchronize()
    iters = 30
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        ops.cache_hit_add(sets[i % 3][0], sets[i % 3][1], out=sets[i % 3][2])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    gbs = n * 10 / (ms * 1e-3) / 1e9
    return {"kernel": "axpb_kernel<bf16,f32,f32> (x + residual, 503.2 MB algorithmic)", "bound": "hbm", "ms": ms, "achieved": gbs, "unit": "GB/s",
            "peak": pk["hbm_gbs"], "frac": gbs / pk["hbm_gbs"], "frac_of_8TBps": gbs / 8000.0, "peak_source": pk["src"],
            "method": "30 back-to-back launches rotating over 3 buffer sets (1.5 GB), CUDA events"}


def cpu_baseline_leg():
    import torch
    state = cpu_sample_setup()
    cores = pick_threads(state)
    tb, ta = [], []
    t0 = time.perf_counter()
    while len(tb) < 5 and time.perf_counter() - t0 < 25:
        b, a = cpu_time_block(state)
        tb.append(b)
        ta.append(a)
    t_block, t_attn = statistics.median(tb), statistics.median(ta)
    value, sec_video, t_miss = cpu_extrapolate(t_block, t_attn, state[-1])
    return {"value": value, "unit": "steps/s", "cores": cores, "kind": "port",
            "sample": (f"{len(tb)} x one WanAttentionBlock (of {LAYERS}) at {state[-1]} of {N_TOK} tokens, torch-CPU oracle, {cores} threads: block {t_block:.3f}s "
                       f"(SDPA {t_attn:.3f}s); attention ~N^2, rest ~N; miss forward {t_miss:.1f}s; video {sec_video:.0f}s")}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=SAMPLE_STEPS)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cache", action="store_true", help="time the non-cached DiT loop at identical shapes")
    ap.add_argument("--skip-cpu", action="store_true", help="omit the cpu_baseline leg (debugging)")
    ap.add_argument("--workload", default="wan1.3b", choices=["wan1.3b", "wan14b"], help="wan14b: BASELINE configs[4] model/shape (not the driver's metric)")
    args = ap.parse_args()
    select_workload(args.workload)
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    if world != args.gpus and args.gpus > 1:
        raise SystemExit(f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE={world})")
    run_ours(args, rank, world)


if __name__ == "__main__":
    main()
