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


class PeerAllReduce:
    """In-kernel SUM over ranks of a flat fp32 buffer through NVLink / NVSwitch peer memory (``nm_grad_allreduce`` in the C ABI,
    ``csrc/nm_collective.cuh``): every rank's kernels write the buffer of step ``e`` into half ``e & 1`` of a symmetric
    allocation (``torch.distributed._symmetric_memory`` is the plumbing: allocation, peer mapping, multicast address), one
    small kernel exchanges arrival flags with the peers and reads the sum with ``multimem.ld_reduce`` -- formed inside the
    switch (NVLS) -- or, without multicast support / with ``NM_AR_NO_MULTICAST=1``, with peer loads in rank order.

        sync = PeerAllReduce(n, device, tail_from=n_params, tail_scale=1 / world)
        buf = sync.half()             # [n] view of this step's half: hand it to the kernels as their output
        ... nm_rollout_forward / nm_rollout_backward write scalars / gradients into buf ...
        sync.all_reduce(out)          # out[n]: sums (entries >= tail_from scaled), identical on every rank
    """

    def __init__(self, n: int, device, group=None, tail_from: Optional[int] = None, tail_scale: float = 1.0):
        import ctypes as C
        import os
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib as L
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self._C, self.lib = C, L.load()
        self.n, self.npad = int(n), (int(n) + 63) // 64 * 64
        self.device = torch.device(device)
        group = group if group is not None else dist.group.WORLD
        floats = int(self.lib.nm_grad_allreduce_buffer_floats(self.n))
        self.buf = symm_mem.empty(floats, dtype=torch.float32, device=self.device)
        self.buf.zero_()
        try:
            self.hdl = symm_mem.rendezvous(self.buf, group)
        except Exception:                                   # older bindings want the group enabled first / take its name
            symm_mem.enable_symm_mem_for_group(group.group_name)
            self.hdl = symm_mem.rendezvous(self.buf, group.group_name)
        self.rank, self.world = int(self.hdl.rank), int(self.hdl.world_size)
        # the tensor may sit at an offset inside the rendezvoused allocation: the same offset on every rank and on the multicast mapping
        delta = self.buf.data_ptr() - int(self.hdl.buffer_ptrs[self.rank])
        self.peers = (C.c_void_p * self.world)(*[int(p) + delta for p in self.hdl.buffer_ptrs])
        mc = int(self.hdl.multicast_ptr or 0)
        mc = mc + delta if mc else 0
        if os.environ.get("NM_AR_NO_MULTICAST") == "1":
            mc = 0
        self.multicast_ptr = mc
        self.mode = "nvls multimem.ld_reduce" if mc else "peer loads in rank order"
        self.tail_from = self.n if tail_from is None else int(tail_from)
        self.tail_scale = float(tail_scale)
        self.epoch = 0
        torch.cuda.synchronize(self.device)
        dist.barrier(group=group)                           # every rank's flags are zero before the first exchange

    def half(self) -> torch.Tensor:
        """``[n]`` view of the half the NEXT ``all_reduce`` call reads."""
        off = ((self.epoch + 1) & 1) * self.npad
        return self.buf[off:off + self.n]

    def all_reduce(self, out: torch.Tensor) -> torch.Tensor:
        C = self._C
        if out.numel() != self.n or out.dtype != torch.float32 or out.device != self.buf.device or not out.is_contiguous():
            raise ValueError("PeerAllReduce.all_reduce: `out` must be a contiguous fp32 tensor of n entries on the buffer's device")
        lo, hi = self.buf.data_ptr(), self.buf.data_ptr() + 4 * self.buf.numel()
        if lo <= out.data_ptr() < hi:
            raise ValueError("PeerAllReduce.all_reduce: `out` must not live in the symmetric buffer (peers are still reading it)")
        self.epoch += 1
        from . import _lib as L
        with torch.cuda.device(self.device):
            L.check(self.lib.nm_grad_allreduce(self.peers, C.c_void_p(self.multicast_ptr or None), C.c_void_p(out.data_ptr()), self.n,
                                               self.rank, self.world, self.epoch, self.tail_from, self.tail_scale,
                                               C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)), "nm_grad_allreduce")
        return out

    def status(self) -> int:
        """0, or 1 when a peer did not arrive within the kernel's time-out (host synchronisation: diagnostics only)."""
        return int(self.buf[2 * self.npad:].view(torch.int32)[17])


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
        # NM_SYNC=peer: CUDA buffers handed to allreduce_flat_ are summed by the in-kernel exchange over NVLink peer memory
        # (PeerAllReduce: one staging copy into the symmetric buffer, then nm_grad_allreduce) instead of NCCL.  Opt-in on this
        # path: the exchange itself is verified on 2 / 4 / 8 GPUs through RawTrainStep (bench.py's default for N > 1,
        # tests/test_peer_allreduce_gpu.py); the Trainer + FusedAdamW route through it has not been run on more than one GPU.
        import os
        self._peer_ok = os.environ.get("NM_SYNC", "nccl") == "peer"
        self._peer = {}

    def shard(self, n_total: int):
        """Row range ``[lo, hi)`` of a global batch owned by this rank (equal shards)."""
        if n_total % self.world_size:
            raise ValueError(f"global batch {n_total} not divisible by world size {self.world_size}")
        per = n_total // self.world_size
        return self.rank * per, (self.rank + 1) * per

    def allreduce_flat_(self, flat: torch.Tensor) -> torch.Tensor:
        """In-place SUM all-reduce of an already flat buffer (the adjoint kernel's output)."""
        if self.world_size > 1:
            if self._peer_ok and flat.is_cuda and flat.dtype == torch.float32 and flat.is_contiguous():
                key = (flat.numel(), flat.device)
                if key not in self._peer:
                    try:
                        self._peer[key] = PeerAllReduce(flat.numel(), flat.device, group=self.group)
                    except Exception:                            # noqa: BLE001 -- no symmetric memory on this box
                        self._peer_ok = False
                sync = self._peer.get(key)
                if sync is not None:
                    sync.half().copy_(flat)
                    return sync.all_reduce(flat)
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
