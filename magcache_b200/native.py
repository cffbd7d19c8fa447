"""`mc_dit_*` from Python: the whole patched Wan forward sequenced by native code (include/magcache_b200.h, csrc/dit_forward.cu).

`NativeWanForward(weights)` hands the packed device weights of a `WanWeights` to `mc_dit_create`; `bind(grid, rope)` gives it one
workspace for a token grid; `forward(...)` is ONE C call per patched forward — prologue, then the hit or the miss branch, then the
head — issuing the launch sequence `WanEngine` issues from Python, through the same entry points (bit-identical results).
`WanEngine(..., native=True)` (or MC_NATIVE=1) routes its plain forwards through it: one sample, one timestep, one GPU, a t2v
model with 16 output channels; everything else stays on the Python-sequenced path. A host that is not Python binds the same
five functions directly.
"""
import ctypes

import torch

from . import _lib

lib, check = _lib.lib, _lib.check


def supported(dims):
    """What csrc/dit_forward.cu sequences: the plain text-to-video model."""
    return dims.model_type == "t2v" and dims.out_dim == 16 and dims.head_dim == 128 and not dims.vace_layers


class NativeWanForward:
    def __init__(self, weights):
        d = weights.dims
        if not supported(d):
            raise NotImplementedError(f"mc_dit_forward sequences t2v models with 16 output channels (got {d.model_type}, out_dim {d.out_dim})")
        self.weights = weights  # keeps the device tensors alive: the handle borrows their pointers
        self.dims = _lib.DitDims(d.dim, d.ffn_dim, d.num_heads, d.num_layers, d.in_dim, d.out_dim, d.freq_dim, d.text_dim, d.text_len, d.eps)
        self._blocks = (_lib.DitBlock * d.num_layers)()
        for i, b in enumerate(weights.blocks):
            for name in _lib.DIT_BLOCK_FIELDS:
                t = b[name]
                assert t.is_contiguous(), name
                setattr(self._blocks[i], name, t.data_ptr())
        self._w = _lib.DitWeights()
        for name in _lib.DIT_TOP_FIELDS:
            t = getattr(weights, name)
            assert t.is_contiguous(), name
            setattr(self._w, name, t.data_ptr())
        self._w.blocks = ctypes.cast(self._blocks, ctypes.POINTER(_lib.DitBlock))
        self.h = lib.mc_dit_create(ctypes.byref(self.dims), ctypes.byref(self._w))
        if not self.h:
            raise _lib.MagCacheError(_lib.MC_ERR_INVALID, lib.mc_last_error().decode("utf-8", "replace"))
        self.grid = None
        self._ws = self._rope = None
        self._launches = {}

    def workspace_bytes(self, grid):
        need = ctypes.c_int64(0)
        check(lib.mc_dit_workspace_bytes(self.h, int(grid[0]), int(grid[1]), int(grid[2]), ctypes.byref(need)))
        return need.value

    def bind(self, grid, rope, device=None):
        """One workspace for the (F, Hp, Wp) token grid; `rope`: fp32 [F*Hp*Wp, 128] cos/sin table (wan.rope_table)."""
        grid = tuple(int(g) for g in grid)
        n = grid[0] * grid[1] * grid[2]
        assert rope.dtype == torch.float32 and rope.is_contiguous() and tuple(rope.shape) == (n, 128)
        need = self.workspace_bytes(grid)
        ws = torch.empty(need + 1024, dtype=torch.uint8, device=rope.device if device is None else device)
        ptr = (ws.data_ptr() + 1023) // 1024 * 1024
        check(lib.mc_dit_bind(self.h, grid[0], grid[1], grid[2], ptr, need, rope.data_ptr()))
        self.grid, self._ws, self._rope, self._ws_ptr = grid, ws, rope, ptr
        self._launches = {}
        return need

    def forward(self, latent, t_dev, context, skip, residual, out=None, stream=None):
        """latent fp32 [C, F, H, W]; t_dev float64 [>= 1] (element 0 is the timestep); context bf16 [text_len, text_dim] zero-padded;
        residual fp32 [N, dim]: read on a hit (`skip`), written on a miss. Returns out fp32 [16, F, H, W]."""
        d, (F, Hp, Wp) = self.weights.dims, self.grid
        n = F * Hp * Wp
        assert latent.dtype == torch.float32 and latent.is_contiguous() and tuple(latent.shape) == (d.in_dim, F, 2 * Hp, 2 * Wp)
        assert t_dev.dtype == torch.float64 and t_dev.is_contiguous() and t_dev.numel() >= 1
        assert residual.dtype == torch.float32 and residual.is_contiguous() and tuple(residual.shape) == (n, d.dim)
        if not skip:
            assert context.dtype == torch.bfloat16 and context.is_contiguous() and tuple(context.shape) == (d.text_len, d.text_dim)
        for t in (latent, t_dev, residual) + (() if skip else (context,)):
            if not t.is_cuda:
                raise RuntimeError("magcache_b200: CUDA tensors only (no CPU path)")
        if out is None:
            out = torch.empty(d.out_dim, F, 2 * Hp, 2 * Wp, dtype=torch.float32, device=latent.device)
        if stream is None:
            stream = torch.cuda.current_stream().cuda_stream
        check(lib.mc_dit_forward(self.h, latent.data_ptr(), t_dev.data_ptr(), None if context is None else context.data_ptr(), int(bool(skip)),
                                 residual.data_ptr(), out.data_ptr(), stream))
        return out

    def plan(self, skip):
        """The launch plan of `forward(skip=...)` as a list of text lines (nothing is launched)."""
        need = ctypes.c_int64(0)
        check(lib.mc_dit_plan(self.h, int(bool(skip)), None, 0, ctypes.byref(need)))
        buf = ctypes.create_string_buffer(need.value)
        check(lib.mc_dit_plan(self.h, int(bool(skip)), buf, need.value, ctypes.byref(need)))
        return buf.value.decode().splitlines()

    def launches(self, skip):
        """Kernel launches one forward issues (plan lines; an attention whose shape splits over the KV range adds its combine kernel)."""
        key = bool(skip)
        if key not in self._launches:
            n = 0
            for line in self.plan(skip):
                n += 1
                if line.startswith("attention "):
                    kv = dict(f.split("=") for f in line.split()[1:])
                    need = ctypes.c_int64(0)
                    check(lib.mc_attn_workspace_bytes(int(kv["Lq"]), int(kv["Lk"]), int(kv["heads"]), ctypes.byref(need)))
                    n += 1 if need.value > 0 else 0
            self._launches[key] = n
        return self._launches[key]

    def close(self):
        if self.h:
            lib.mc_dit_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass
