"""
``FusedAdamW``: gradient clipping + AdamW as ONE call into the C ABI (``nm_adamw_step``) over the flat parameter arena
the rollout kernels read and the flat gradient buffer the adjoint kernels wrote (replaces the tail of
the reference's training step, trainer.py:251-252 ``clip_grad_norm_`` + ``optimizer.step()``, for a fused problem).

The arena IS the parameter storage (module parameters are views of it, ``rollout.ParamArena``), so the update needs
no gather / scatter: two small launches on the current stream, no host synchronisation, arithmetic of
``torch.optim.AdamW`` (decoupled weight decay, bias-corrected moments).  Data-parallel: all-reduce the flat gradient
first (``GradientSync.allreduce_flat_``), then step -- every rank applies the same update.

Covers the parameters that live in the arena of ONE fused rollout plan.  Anything else the problem trains (parameters
of structured slim maps, of python-evaluated terms, of a second system) is not in the arena: the constructor refuses
such a problem instead of silently leaving those parameters untrained -- use a torch optimiser there.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L
from .slim.linear import DerivedWeight


class FusedAdamW:
    def __init__(self, problem, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_grad_norm=0.0):
        self.problem = problem
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)
        self.max_grad_norm = float(max_grad_norm)
        self.step_count = 0
        self.last_grad_norm = None
        self._runner = None
        self._m = self._v = self._mask = self._ws = self._norm = None
        self.param_groups = [{"lr": self.lr}]          # what lr schedulers look at

    # -- binding to the plan's runner (exists after the first forward call)
    def _bind(self):
        systems = [n for n in self.problem.nodes if hasattr(n, "_plans")]
        runners = [getattr(p, "_runner", None) for s in systems for p in s._plans.values()]
        runners = [r for r in runners if r is not None]
        if len(runners) != 1:
            raise RuntimeError(f"FusedAdamW needs exactly one fused rollout plan that has run a forward pass, found {len(runners)}")
        r = runners[0]
        plan = r.plan
        if any(isinstance(q, DerivedWeight) for q in plan.params):
            raise RuntimeError("FusedAdamW: structured slim maps keep their own parameters outside the arena -- use a torch optimiser")
        in_arena = {id(q) for q in plan.params}
        stray = [n for n, q in self.problem.named_parameters() if q.requires_grad and id(q) not in in_arena]
        if stray:
            raise RuntimeError(f"FusedAdamW: trainable parameters outside the fused plan's arena: {stray}")
        flat = r.arena.flat
        self._runner = r
        self._m, self._v = torch.zeros_like(flat), torch.zeros_like(flat)
        mask = torch.zeros(flat.numel(), dtype=torch.uint8, device=flat.device)
        for q, (off, n) in zip(plan.params, plan.param_slices):
            if q.requires_grad:
                mask[off:off + n] = 1
        self._mask = mask
        nbytes = L.load().nm_adamw_workspace_bytes(flat.numel())
        self._ws = torch.empty(nbytes, dtype=torch.uint8, device=flat.device)
        self._norm = torch.zeros(1, dtype=torch.float32, device=flat.device)

    def zero_grad(self, set_to_none=True):
        for q in self.problem.parameters():
            q.grad = None

    def step(self, grad_flat=None):
        """One update from ``grad_flat`` (default: the flat gradient of the last backward pass, ``runner.last_grad_flat``
        -- all-reduce it in place first when training data-parallel)."""
        if self._runner is None:
            self._bind()
        r = self._runner
        if not r.arena.in_sync(r.plan.params):
            r.arena.adopt(r.plan.params)
        g = r.last_grad_flat if grad_flat is None else grad_flat
        if g is None:
            raise RuntimeError("FusedAdamW.step: no gradient -- call loss.backward() first")
        flat = r.arena.flat
        self.step_count += 1
        self.lr = float(self.param_groups[0]["lr"])
        lib = L.load()
        stream = C.c_void_p(torch.cuda.current_stream(flat.device).cuda_stream)
        with torch.cuda.device(flat.device):
            L.check(lib.nm_adamw_step(flat.data_ptr(), g.data_ptr(), self._m.data_ptr(), self._v.data_ptr(), self._mask.data_ptr(),
                                      flat.numel(), self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay,
                                      self.step_count, self.max_grad_norm, self._norm.data_ptr(), self._ws.data_ptr(),
                                      self._ws.numel(), stream), "nm_adamw_step")
        self.last_grad_norm = self._norm          # device scalar (no host sync)

    def state_dict(self):
        return {"step": self.step_count, "exp_avg": None if self._m is None else self._m.clone(),
                "exp_avg_sq": None if self._v is None else self._v.clone(), "lr": self.lr}

    def load_state_dict(self, sd):
        self.step_count, self.lr = int(sd["step"]), float(sd["lr"])
        self.param_groups[0]["lr"] = self.lr
        if sd["exp_avg"] is not None:
            if self._runner is None:
                self._bind()
            self._m.copy_(sd["exp_avg"]); self._v.copy_(sd["exp_avg_sq"])
