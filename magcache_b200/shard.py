"""Token-axis sharding of the Wan forward over the GPUs of one node.

Everything in the forward is token-local except self-attention, which needs every key/value. Each rank owns a contiguous
range of tokens (contiguous in (f, h, w) raster order — the "temporal/token shard"), keeps its slice of the residual stream
and of the residual cache (as the reference's only sequence-parallel MagCache does, eval/magcache/experiments/opensora.py:310,347),
and per layer contributes its K and V rows to an all-gather — the B200 counterpart of videosys/core/comm.py:282-292
(`dist.all_gather` + `torch.cat`): here `all_gather_into_tensor` writes straight into the attention kernel's K buffer,
no list of tensors and no concatenation copy. The controller is a pure function of `cnt` and the table, so every rank takes the
same hit/miss decision without communicating.

The functions below are plain torch / torch.distributed plumbing (they run on CPU tensors with gloo in the test-suite and on
CUDA tensors with NCCL in the engine).
"""
from dataclasses import dataclass

import torch.distributed as dist


@dataclass
class TokenShard:
    rank: int
    world: int
    n_tokens: int
    group: object = None

    def __post_init__(self):
        if self.n_tokens % self.world != 0:
            # the pad rule of videosys/core/comm.py:373-378 is not needed for the shipped shapes
            # (32760 = 8*4095, 75600 = 8*9450, 118800 = 8*14850); refuse instead of silently mis-sharding
            raise ValueError(f"token count {self.n_tokens} is not divisible by world size {self.world}")

    @property
    def n_local(self):
        return self.n_tokens // self.world

    @property
    def start(self):
        return self.rank * self.n_local

    @property
    def stop(self):
        return self.start + self.n_local

    def rows(self, t):
        """This rank's rows of a token-major tensor [N, ...]."""
        return t[self.start:self.stop]


def gather_rows(local, full, group=None, async_op=False):
    """All-gather token rows: full[r*n_local:(r+1)*n_local] = rank r's `local` ([n_local, C], contiguous)."""
    assert local.is_contiguous() and full.is_contiguous() and full.shape[0] % local.shape[0] == 0
    return dist.all_gather_into_tensor(full, local, group=group, async_op=async_op)


def sum_partial_outputs(out, group=None):
    """Every rank wrote only its tokens' positions of the (zero-initialised) output: the sum over ranks is the full tensor,
    exactly (x + 0 == x), and leaves it replicated for the caller's scheduler step."""
    dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out


def allreduce_stats(stats, group=None):
    """Calibration statistics of a sharded run: (sum ratio, sum ratio^2, sum (1-cos), rows) add across ranks."""
    dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats
