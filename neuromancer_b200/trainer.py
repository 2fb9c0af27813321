"""
``Trainer``: the caller of the hot path (reference trainer.py:154-344).

Same loop -- ``zero_grad -> output[train_metric].backward() -> clip_grad_norm_ -> optimizer.step()``
with best-model tracking on the dev metric -- with an optional ``GradientSync`` that all-reduces the
gradients between backward and the clip when training data-parallel.  Logging / callbacks /
Lightning integration of the reference are outside the hot-path scope.
"""
from __future__ import annotations

from copy import deepcopy

import numpy as np
import torch

from .dataset import DevicePrefetcher, move_batch_to_device
from .optim import FusedAdamW


class LaggedScalarReader:
    """Per-step device->host read of a scalar (the loss) that does not stall the launch pipeline: step k's value is
    copied into pinned host memory asynchronously and collected while step k+1 is already queued.  Every step's
    value reaches the host; `read()` returns the values in order (one step late), `drain()` the rest."""

    def __init__(self, device, depth=2):
        self.buf = [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(depth)]
        self.ev = [torch.cuda.Event() for _ in range(depth)]
        self.pending, self.k, self.depth = [], 0, depth

    def push(self, scalar):
        out = []
        if len(self.pending) == self.depth:
            out.append(self._pop())
        slot = self.k % self.depth
        self.buf[slot].copy_(scalar.detach().reshape(1), non_blocking=True)
        self.ev[slot].record()
        self.pending.append(slot)
        self.k += 1
        return out

    def _pop(self):
        slot = self.pending.pop(0)
        self.ev[slot].synchronize()
        return float(self.buf[slot])

    def drain(self):
        return [self._pop() for _ in range(len(self.pending))]


class Trainer:
    def __init__(self, problem, train_data, dev_data=None, test_data=None, optimizer=None, logger=None,
                 callback=None, lr_scheduler=False, epochs=1000, epoch_verbose=1, patience=5, warmup=0,
                 train_metric="train_loss", dev_metric="dev_loss", test_metric="test_loss",
                 eval_metric="dev_loss", eval_mode="min", clip=100.0, multi_fidelity=False,
                 device="cuda", grad_sync=None):
        self.model = problem
        self.optimizer = optimizer if optimizer is not None else torch.optim.Adam(problem.parameters(), 0.01, betas=(0.0, 0.9))
        self.train_data, self.dev_data, self.test_data = train_data, dev_data, test_data
        self.epochs, self.current_epoch, self.epoch_verbose = epochs, 0, epoch_verbose
        self.train_metric, self.dev_metric, self.test_metric = train_metric, dev_metric, test_metric
        self.eval_metric, self._eval_min = eval_metric, eval_mode == "min"
        self.patience, self.warmup, self.badcount, self.clip = patience, warmup, 0, clip
        self.best_devloss = np.finfo(np.float32).max if self._eval_min else 0.
        self.best_model = deepcopy(self.model.state_dict())
        self.device, self.grad_sync = device, grad_sync
        if lr_scheduler and isinstance(self.optimizer, FusedAdamW):
            raise ValueError("lr_scheduler=True needs a torch optimiser (ReduceLROnPlateau); set FusedAdamW.param_groups[0]['lr'] yourself")
        self.lr_scheduler = (torch.optim.lr_scheduler.ReduceLROnPlateau(self.optimizer, mode="min", factor=0.5, patience=100)
                             if lr_scheduler else None)

    def train_step(self, batch):
        output = self.model(batch)
        self.optimizer.zero_grad()
        output[self.train_metric].backward()
        if isinstance(self.optimizer, FusedAdamW):
            # clip + AdamW in one call over the flat gradient the adjoint kernels wrote (all-reduced in place first)
            if self.optimizer._runner is None:
                self.optimizer._bind()
            if self.grad_sync is not None:
                self.grad_sync.allreduce_flat_(self.optimizer._runner.last_grad_flat)
            self.optimizer.max_grad_norm = float(self.clip) if self.clip else 0.0
            self.optimizer.step()
            return output
        if self.grad_sync is not None:
            self.grad_sync.sync_gradients(self.model.parameters())
        torch.nn.utils.clip_grad_norm_(self.model.parameters(), self.clip)
        self.optimizer.step()
        return output

    def train(self):
        output = {}
        for i in range(self.current_epoch, self.current_epoch + self.epochs):
            self.model.train()
            losses = []
            for batch in DevicePrefetcher(self.train_data, self.device):     # copy of batch k+1 overlaps step k
                batch["epoch"] = i
                output = self.train_step(batch)
                losses.append(output[self.train_metric].detach())
            output[f"mean_{self.train_metric}"] = torch.mean(torch.stack(losses))
            if self.lr_scheduler is not None:
                self.lr_scheduler.step(output[f"mean_{self.train_metric}"])
            with torch.set_grad_enabled(self.model.grad_inference):
                self.model.eval()
                if self.dev_data is not None:
                    losses = []
                    for batch in self.dev_data:
                        ev = self.model(move_batch_to_device(batch, self.device))
                        losses.append(ev[self.dev_metric].detach())
                    output[self.dev_metric] = ev[self.dev_metric]
                    output[f"mean_{self.dev_metric}"] = torch.mean(torch.stack(losses))
                score = output.get(self.eval_metric, output[f"mean_{self.train_metric}"])
                better = score < self.best_devloss if self._eval_min else score > self.best_devloss
                if better:
                    self.best_model = deepcopy(self.model.state_dict())
                    self.best_devloss = score
                    self.badcount = 0
                elif i > self.warmup:
                    self.badcount += 1
                if self.epoch_verbose and i % self.epoch_verbose == 0:
                    print(f"epoch: {i}  {self.train_metric}: {float(output[f'mean_{self.train_metric}'])}")
                if self.badcount > self.patience:
                    print("Early stopping!!!")
                    break
                self.current_epoch = i + 1
        self.model.load_state_dict(self.best_model)
        return self.best_model
