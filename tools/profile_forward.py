"""One Wan2.1-1.3B-shaped forward with a reduced layer count (default 1) at the full 32760-token shape — the ncu target.
   ncu --set full -k regex:attn_fwd -c 1 python tools/profile_forward.py"""
import sys

import torch

sys.path.insert(0, ".")
import magcache_b200 as mc  # noqa: E402

layers = int(sys.argv[1]) if len(sys.argv) > 1 else 1
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda", 0)
dims = mc.WanDims(1536, 8960, 12, layers)
model = mc.WanModelHandle(mc.WanWeights.random(dims, dev, seed=0))
mc.init_magcache(model, 50, thresh=0.12, K=4, retention_ratio=0.2, table="wan2.1_t2v_1.3b")
lat = torch.randn(16, 21, 60, 104, device=dev)
ctx = torch.randn(512, 4096, device=dev).bfloat16()
t = torch.tensor([900.0], device=dev)
for _ in range(reps):
    out = model([lat], t=t, context=[ctx], seq_len=32760)[0]
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
