"""K1 cache-hit add at the Wan2.1-1.3B shape on rotating buffers — the ncu target for the HBM-bound kernel."""
import sys

import torch

sys.path.insert(0, ".")
from magcache_b200 import ops  # noqa: E402

n = 32760 * 1536
sets = [(torch.randn(n, device="cuda").bfloat16(), torch.randn(n, device="cuda"), torch.empty(n, device="cuda")) for _ in range(3)]
for i in range(6):
    ops.cache_hit_add(sets[i % 3][0], sets[i % 3][1], out=sets[i % 3][2])
torch.cuda.synchronize()
print("ok")
