"""
``Problem``: computational graph + loss.

API mirror of the reference's ``problem.py`` (``Problem`` :148-210): ``problem(batch)`` returns a
dictionary whose keys are prefixed with ``batch['name']`` and contain the trajectories, every loss
term, ``objective_loss``, ``penalty_loss`` and ``loss``.

When the node list is a single ``System`` and the loss a ``PenaltyLoss`` -- the shape of every DPC
and system-ID example -- rollout AND scoring run in the fused sm_100a kernels and ``.backward()`` on
any returned scalar runs the fused adjoint kernel.  The per-constraint ``*_value`` / ``*_violation``
tensors and the ``C_*`` matrices are materialised lazily on first access (they double the HBM
traffic of the step and nothing in the training loop reads them).
"""
from __future__ import annotations

from typing import Callable, Dict, List

import torch
import torch.nn as nn

from .loss import AggregateLoss, BarrierLoss, PenaltyLoss
from .system import System


class LazyOutputs(dict):
    """dict whose missing keys may be produced on demand by registered thunks."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._thunks = {}

    def register(self, keys, thunk):
        """``thunk()`` returns a dict containing (at least) ``keys``."""
        for k in keys:
            self._thunks[k] = thunk

    def __missing__(self, key):
        thunk = self._thunks.get(key)
        if thunk is None:
            raise KeyError(key)
        produced = thunk()
        for k, v in produced.items():
            if not dict.__contains__(self, k):
                dict.__setitem__(self, k, v)
            self._thunks.pop(k, None)
        return dict.__getitem__(self, key)

    def __contains__(self, key):
        return dict.__contains__(self, key) or key in self._thunks

    def lazy_keys(self):
        return list(self._thunks)

    def materialise(self):
        for k in list(self._thunks):
            if k in self._thunks:
                self[k]
        return self


class Problem(nn.Module):
    def __init__(self, nodes: List[Callable], loss: Callable, grad_inference=False, check_overwrite=False):
        super().__init__()
        self.nodes = nn.ModuleList(nodes)
        self.loss = loss
        self.grad_inference = grad_inference
        self.check_overwrite = check_overwrite
        self._check_keys()

    def _check_keys(self):
        import warnings
        seen = set()
        for node in [*self.nodes, self.loss]:
            seen |= set(node.input_keys)
            clash = set(node.output_keys) & seen
            if self.check_overwrite and clash:
                warnings.warn(f"Keys {clash} are being overwritten by the node {node}.")
            seen |= set(node.output_keys)

    def _fusable(self):
        # exactly PenaltyLoss, or exactly BarrierLoss with one of its named barrier functions (lowered with the barrier kind
        # in the term descriptors); other subclasses redefine how constraints enter the loss and are evaluated with torch
        # ops on the trajectories the fused rollout wrote
        if len(self.nodes) != 1 or not isinstance(self.nodes[0], System):
            return False
        return type(self.loss) is PenaltyLoss or (type(self.loss) is BarrierLoss and self.loss.barrier_name is not None)

    def step(self, input_dict: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """Run the nodes (each a fused ``System``) and merge their outputs."""
        for node in self.nodes:
            out = node(input_dict)
            if isinstance(out, torch.Tensor):
                out = {node.name: out}
            input_dict = {**input_dict, **out}
        return input_dict

    def forward(self, data: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        if self._fusable():
            from .rollout import run_problem
            out = run_problem(self.nodes[0], self.loss, data)
        else:
            out = None
        if out is None:          # not fusable, or a BarrierLoss constraint that could not be lowered
            out = self.loss(self.step(data))
            if isinstance(out, torch.Tensor):
                out = {self.loss.name: out}
        prefix = data["name"]
        res = LazyOutputs({f"{prefix}_{k}": v for k, v in dict.items(out)})
        if isinstance(out, LazyOutputs):
            for k in out.lazy_keys():
                res.register([f"{prefix}_{k}"],
                             (lambda k=k: {f"{prefix}_{k}": out[k]}))
        return res

    def graph(self, include_objectives=True):
        raise NotImplementedError("graph drawing (pydot) is outside the hot-path scope")

    def show(self, figname=None):
        raise NotImplementedError("graph drawing (pydot) is outside the hot-path scope")

    def __repr__(self):
        s = "### MODEL SUMMARY ###\n\nCOMPONENTS:"
        for c in self.nodes:
            s += f"\n  {repr(c)}"
        s += "\n\nCONSTRAINTS:" if len(self.loss.constraints) else "\n\nCONSTRAINTS: none"
        for c in self.loss.constraints:
            s += f"\n  {repr(c)}"
        s += "\n\nOBJECTIVES:" if len(self.loss.objectives) else "\n\nOBJECTIVES: none"
        for c in self.loss.objectives:
            s += f"\n  {repr(c)}"
        return s
