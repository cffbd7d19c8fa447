"""Bring-up aid for the MN-major V descriptor of the attention kernels: runs a small attention through both kernels with the
descriptor strides as built (MC_ATTN_VDESC=0) and swapped (=1) and prints the error against an fp64 reference."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from magcache_b200 import ops  # noqa: E402

dev = torch.device("cuda")
torch.manual_seed(0)
for Lq, Lk, heads in ((128, 128, 1), (256, 384, 2), (300, 1300, 2)):
    W = heads * 128
    q, k, v = (torch.randn(n, W, device=dev).bfloat16() for n in (Lq, Lk, Lk))
    qh, kh, vh = (t.double().view(-1, heads, 128).transpose(0, 1) for t in (q, k, v))
    ref = (torch.softmax(qh @ kh.transpose(1, 2) / math.sqrt(128), -1) @ vh).transpose(0, 1).reshape(Lq, W).float()
    for kern in ("1", "2"):
        for mode in ("0", "1"):
            os.environ["MC_ATTN_KERNEL"], os.environ["MC_ATTN_VDESC"] = kern, mode
            try:
                out = ops.attention(q, k, v, heads)
                torch.cuda.synchronize()
                err = float((out.float() - ref).abs().max())
            except Exception as ex:  # noqa: BLE001
                err = f"failed: {ex}"
            print(f"Lq={Lq} Lk={Lk} heads={heads} kernel={'short' if kern == '1' else 'long'} vdesc={mode}: max abs err {err}", flush=True)
os.environ.pop("MC_ATTN_KERNEL", None)
os.environ.pop("MC_ATTN_VDESC", None)
