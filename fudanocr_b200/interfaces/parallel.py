"""Data-parallel plumbing that replaces the reference's ``nn.DataParallel`` wrap
(scene-text-telescope/interfaces/base.py:175-187, loss/text_focus_loss.py:57).

The reference runs ONE process that scatters the batch, re-broadcasts all weights every step and reduces
gradients onto GPU 0.  Here: one process per GPU (torchrun), weights replicated once, each rank takes the row
shard ``[r*B/N, (r+1)*B/N)`` of every batch tensor (BatchNorm statistics stay per replica exactly as under
DataParallel), and the only exchange is one all-reduce of the flat gradient buffer.  The
gradient is averaged (the losses are means), then clipped (global L2 of the *averaged* gradient, so every rank
takes the identical step) — see ``TBSRNTrainer``.

``DataParallel`` keeps the attribute surface the reference loops touch: ``.module``, ``.parameters()``,
``.train()/.eval()``, ``__call__``, ``state_dict()/load_state_dict()`` (base.py:179-187, 261-266).
"""
from __future__ import annotations

from typing import Sequence, Tuple

import torch
import torch.distributed as dist
from torch import nn


def shard_bounds(n_rows: int, rank: int, world: int) -> Tuple[int, int]:
    """rows [lo, hi) of a global batch owned by `rank` (contiguous, sizes differ by at most one)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, rem = divmod(n_rows, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(tensors: Sequence, rank: int, world: int):
    """slice every batch-major tensor / list (e.g. label strings) to this rank's rows"""
    n = len(tensors[0])
    lo, hi = shard_bounds(n, rank, world)
    return [t[lo:hi] for t in tensors]


def allreduce_mean_(flat_grad: torch.Tensor, group=None) -> torch.Tensor:
    """in-place mean over ranks of one flat gradient buffer (NCCL on GPUs, gloo in the CPU tests)"""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
        flat_grad.div_(dist.get_world_size(group))
    return flat_grad


class DataParallel(nn.Module):
    """Process-per-GPU stand-in for ``nn.DataParallel(model, device_ids=range(ngpu))``: wraps the local replica.
    The reference's ``save_checkpoint`` dereferences ``.module`` unconditionally (base.py:261,266), so this wrapper
    is used even on one GPU."""

    def __init__(self, module: nn.Module, device_ids=None):
        super().__init__()
        self.module = module

    def forward(self, *a, **k):
        return self.module(*a, **k)
