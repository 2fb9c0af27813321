"""
Data-parallel training of a fused ``Problem``: one process per GPU, the sampled batch sharded
across ranks, ONE all-reduce of the flat gradient (+ the loss scalars) per step.

The reference has no communication layer of its own; multi-GPU training is delegated to Lightning
DDP (reference trainer.py:49,112-116), whose semantics are: each rank computes the *local-mean* loss
on its shard and gradients are *averaged*.  Here the adjoint kernel already divides by the global
trajectory count (``b_total = B_local * world_size``), so a SUM all-reduce yields the same global-mean
gradient; the loss scalars are averaged.  The payload is the few-KB flat gradient the adjoint kernel
wrote, so the collective is latency-bound (NCCL over NVLink; gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist

from . import rollout


class GradientSync:
    """All-reduce helper bound to a process group."""

    def __init__(self, group=None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.group = group
        self.world_size = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        rollout.set_world_size(self.world_size)
        self._flat: Optional[torch.Tensor] = None

    def shard(self, n_total: int):
        """Row range ``[lo, hi)`` of a global batch owned by this rank (equal shards)."""
        if n_total % self.world_size:
            raise ValueError(f"global batch {n_total} not divisible by world size {self.world_size}")
        per = n_total // self.world_size
        return self.rank * per, (self.rank + 1) * per

    def allreduce_flat_(self, flat: torch.Tensor) -> torch.Tensor:
        """In-place SUM all-reduce of an already flat buffer (the adjoint kernel's output)."""
        if self.world_size > 1:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        return flat

    def sync_gradients(self, params: Iterable[torch.nn.Parameter], scalars: Optional[torch.Tensor] = None):
        """Sum ``p.grad`` over ranks with a single collective; optionally average ``scalars``
        (loss values) in the same message.  Returns the averaged scalars (or None)."""
        params = [p for p in params if p.grad is not None]
        if self.world_size == 1:
            return scalars
        pieces: List[torch.Tensor] = [p.grad.reshape(-1) for p in params]
        n_s = 0
        if scalars is not None:
            s = scalars.detach().reshape(-1).to(pieces[0].dtype if pieces else scalars.dtype)
            pieces.append(s / self.world_size)
            n_s = s.numel()
        flat = torch.cat(pieces)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        off = 0
        for p in params:
            n = p.grad.numel()
            p.grad.copy_(flat[off:off + n].view_as(p.grad))
            off += n
        return flat[off:off + n_s].view(scalars.shape) if n_s else None

    def close(self):
        rollout.set_world_size(1)
