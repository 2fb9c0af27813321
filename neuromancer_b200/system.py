"""
``Node`` / ``System`` / ``SystemPreview``: the closed-loop rollout.

API mirror of the reference's ``system.py`` (``Node`` :25-76, ``System`` :115-291): same constructor
arguments, ``input_keys`` / ``output_keys``, ``nsteps`` / ``nstep_key`` / ``init_func`` semantics and
the same output dictionary (all inputs plus every rolled-out key as ``[batch, time, dim]``).

The execution is not a port.  The reference runs ``for i in range(nsteps): for node in nodes`` in
Python, slicing ``[:, i]`` and re-concatenating the whole history each step (system.py:272-276,
234-246).  Here ``forward`` compiles the node list once into an ``NmPlan`` (``plan.py``) and runs the
whole horizon for the whole batch in ONE persistent sm_100a kernel (``csrc/``), writing each
trajectory entry exactly once; the backward-through-time pass is a matching fused adjoint kernel.
There is no eager / CPU fallback: tensors must live on a CUDA device and the node list must be
recognisable, otherwise ``forward`` raises.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .constraint import Variable
from . import plan as _plan


class Node(nn.Module):
    """Wraps a callable with the dictionary keys it reads and writes (reference system.py:25-76)."""

    def __init__(self, callable, input_keys, output_keys, name=None):
        super().__init__()
        self.input_keys = [k.key if isinstance(k, Variable) else k for k in input_keys]
        self.output_keys = [k.key if isinstance(k, Variable) else k for k in output_keys]
        self.callable, self.name = callable, name

    def forward(self, data):
        """Single evaluation on a dict of ``[batch, dim]`` tensors (outside a fused rollout)."""
        result = self.callable(*[data[k] for k in self.input_keys])
        if not isinstance(result, tuple):
            result = (result,)
        return dict(zip(self.output_keys, result))

    def freeze(self):
        for p in self.callable.parameters():
            p.requires_grad = False

    def unfreeze(self):
        for p in self.callable.parameters():
            p.requires_grad = True

    def __repr__(self):
        return f"{self.name}({', '.join(self.input_keys)}) -> {', '.join(self.output_keys)}"


class System(nn.Module):
    """Cyclic computation over a horizon, executed by the fused rollout kernels."""

    def __init__(self, nodes, name=None, nstep_key="X", init_func=None, nsteps=None):
        super().__init__()
        self.nstep_key, self.nsteps = nstep_key, nsteps
        self.nodes, self.name = nn.ModuleList(nodes), name
        if init_func is not None:
            self.init = init_func
        self._name_nodes()
        produced, needed = [], []
        for node in self.nodes:
            needed += [k for k in node.input_keys if k not in produced]
            produced += node.output_keys
        self.input_keys = list(dict.fromkeys(needed))
        self.output_keys = list(dict.fromkeys(produced))
        self._layout = None
        self._plans = {}

    def _name_nodes(self):
        n = 1
        for node in self.nodes:
            if node.name is None or node.name == "":
                node.name = f"node_{n}"
                n += 1
        names = [node.name for node in self.nodes]
        assert len(set(names)) == len(names), \
            "All system nodes must have unique names to construct a computational graph."

    def init(self, data):
        """Hook to set initial conditions; identity by default (reference system.py:248-259)."""
        return data

    # -- lowering -------------------------------------------------------------------------
    def _preview_spec(self):
        """``(preview_keys_map, preview_length, pad_mode, pad_constant)`` -- None for a plain System."""
        return None

    def layout(self) -> _plan.SystemLayout:
        if self._layout is None:
            self._layout = _plan.recognise(list(self.nodes), preview=self._preview_spec())
        return self._layout

    def resolve_nsteps(self, data):
        if self.nsteps is not None:
            return int(self.nsteps)
        return int(data[self.nstep_key].shape[1])

    def plan_for(self, data, loss=None) -> _plan.RolloutPlan:
        """The (cached) ``RolloutPlan`` for these tensor shapes."""
        layout = self.layout()
        for k in self.input_keys:
            if k not in data:
                raise KeyError(k)
        nsteps = self.resolve_nsteps(data)
        shapes = {k: tuple(v.shape) for k, v in data.items() if isinstance(v, torch.Tensor)}
        key = (nsteps, id(loss), tuple(sorted((k, s[1:]) for k, s in shapes.items())))
        plan = self._plans.get(key)
        if plan is not None and plan.values_stale():     # a matrix / constant / bound the plan carries by value has changed
            plan = None
        if plan is None:
            plan = _plan.RolloutPlan(layout, nsteps, shapes, loss=loss)
            self._plans = {key: plan}        # keep one: shapes rarely change during training
        return plan

    def forward(self, input_dict):
        """Roll the system out for ``nsteps`` and return the input dict extended with the
        trajectories (``x`` as ``[B, nsteps+1, nx]``, actions / outputs as ``[B, nsteps, d]``)."""
        from .rollout import run_rollout
        data = self.init(dict(input_dict))
        plan = self.plan_for(data)
        out, _ = run_rollout(plan, data)
        return out

    def freeze(self):
        for node in self.nodes:
            node.freeze()

    def unfreeze(self):
        for node in self.nodes:
            node.unfreeze()

    def graph(self):
        raise NotImplementedError("graph drawing (pydot) is outside the hot-path scope")

    def show(self, figname=None):
        raise NotImplementedError("graph drawing (pydot) is outside the hot-path scope")


class SystemPreview(System):
    """``System`` whose nodes may read a *window* of future values of known signals
    (reference system.py:293-367): for every key ``k`` in ``preview_keys_map`` and every node named in
    ``preview_keys_map[k]`` the node receives ``data[k][:, i : i+1+L_k, :]`` flattened time-major to
    ``[batch, (L_k+1)*d]``; where the window runs past the signal it is right-padded the way
    ``torch.nn.functional.pad(mode=pad_mode)`` pads the whole signal (replicate / circular / constant /
    reflect).  The fused rollout gathers the window inside the kernel (pure indexing, bit-exact) instead
    of materialising look-ahead tensors, which cuts the exogenous HBM traffic by the window length."""

    def __init__(self, nodes, preview_keys_map: dict = {}, preview_length: dict = None, pad_mode: str = "circular",
                 pad_constant=0.0, name=None, nstep_key="X", init_func=None, nsteps=None):
        super().__init__(nodes=nodes, name=name, nstep_key=nstep_key, init_func=init_func, nsteps=nsteps)
        self.preview_keys_map = preview_keys_map
        self.pad_mode = pad_mode
        self.pad_constant = pad_constant if pad_mode == "constant" else None
        self.preview_length = (preview_length if preview_length is not None
                               else {k: nsteps for k in preview_keys_map})

    def _preview_spec(self):
        return self.preview_keys_map, self.preview_length, self.pad_mode, self.pad_constant
