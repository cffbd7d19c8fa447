"""Bring-up diagnostics for the tcgen05 kernels (run on the GPU box). Prints structured mismatch information instead of
asserting, so one gpurun call tells us *how* a descriptor / swizzle / barrier is wrong."""
import math
import sys

import torch

sys.path.insert(0, ".")
from magcache_b200 import _lib, ops  # noqa: E402

torch.manual_seed(0)
dev = "cuda"


def report(name, got, ref, tol=2e-2):
    err = (got.float() - ref.float()).abs()
    bad = err > tol * (1 + ref.float().abs())
    print(f"[{name}] shape={tuple(got.shape)} max_err={err.max().item():.4g} mean_err={err.mean().item():.4g} bad={int(bad.sum())}/{bad.numel()}", flush=True)
    if bad.any():
        rows = bad.any(1).nonzero().flatten()
        cols = bad.any(0).nonzero().flatten()
        print("   bad rows (first 16):", rows[:16].tolist(), "... count", rows.numel())
        print("   bad cols (first 16):", cols[:16].tolist(), "... count", cols.numel())
        r0 = int(rows[0])
        print("   row", r0, "got", got[r0, :8].float().tolist())
        print("   row", r0, "ref", ref[r0, :8].float().tolist())
    return not bad.any()


def gemm_case(M, N, K, epi=_lib.MC_EPI_BIAS_F32):
    a = torch.randn(M, K, device=dev).bfloat16()
    b = (torch.randn(N, K, device=dev) / math.sqrt(K)).bfloat16()
    out = ops.gemm(a, b, None, epi)
    torch.cuda.synchronize()
    return report(f"gemm {M}x{N}x{K}", out, a.float() @ b.float().t())


def gemm_onehot():
    # A one-hot in k -> out[m, n] = B[n, k(m)] : decodes which k-slice every row really reads
    M, N, K = 128, 128, 64
    a = torch.zeros(M, K, device=dev)
    km = torch.arange(M, device=dev) % K
    a[torch.arange(M, device=dev), km] = 1
    b = torch.arange(N * K, device=dev, dtype=torch.float32).reshape(N, K) % 251
    out = ops.gemm(a.bfloat16(), b.bfloat16(), None, _lib.MC_EPI_BIAS_F32)
    torch.cuda.synchronize()
    ok = report("gemm onehot", out, a @ b.bfloat16().float().t(), tol=1e-3)
    if not ok:
        bb = b.bfloat16().float()
        for m in [0, 1, 2, 8, 9, 17, 64, 127]:
            hits = [(int(k)) for k in range(K) if torch.allclose(out[m], bb[:, k])]
            print(f"   row {m}: expected k={int(km[m])}, matches k in {hits}")


def attn_case(Lq, Lk, heads):
    W = heads * 128
    q = torch.randn(Lq, W, device=dev).bfloat16()
    k = torch.randn(Lk, W, device=dev).bfloat16()
    v = torch.randn(Lk, W, device=dev).bfloat16()
    ld = (Lk + 7) // 8 * 8
    vt = torch.zeros(W, ld, dtype=torch.bfloat16, device=dev)
    vt[:, :Lk] = v.t()
    out = ops.attention(q, k, vt[:, :Lk], heads)
    torch.cuda.synchronize()
    qh = q.float().view(Lq, heads, 128).transpose(0, 1)
    kh = k.float().view(Lk, heads, 128).transpose(0, 1)
    vh = v.float().view(Lk, heads, 128).transpose(0, 1)
    ref = (torch.softmax(qh @ kh.transpose(1, 2) / math.sqrt(128), -1) @ vh).transpose(0, 1).reshape(Lq, W)
    return report(f"attn Lq={Lq} Lk={Lk} H={heads}", out, ref, tol=3e-2)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0), flush=True)
    if which in ("all", "gemm"):
        gemm_onehot()
        for shp in [(128, 128, 64), (128, 128, 128), (128, 128, 512), (256, 256, 1536), (300, 200, 512), (1000, 1536, 1536)]:
            gemm_case(*shp)
    if which in ("all", "attn"):
        for shp in [(128, 64, 1), (128, 128, 1), (128, 256, 1), (256, 512, 2), (300, 1000, 3), (1000, 4095, 12)]:
            attn_case(*shp)
