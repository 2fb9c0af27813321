"""
Loss aggregators.  ``PenaltyLoss`` is the one on the hot path.

API mirror of the reference (``loss.py``: ``AggregateLoss`` :25-152, ``PenaltyLoss`` :155-181):
``loss = sum(objectives) + sum(constraints)`` in list order, plus the per-constraint value /
violation matrices ``C_values``, ``C_violations`` and their equality / inequality column splits.

Inside a ``Problem`` whose node is a ``System`` the terms are lowered to ``NmTerm`` descriptors and
scored by the fused kernels (see ``neuromancer_b200.plan``); ``forward`` below is the eager
evaluation used for terms that cannot be lowered and for the lazily materialised ``C_*`` outputs.
"""
import math

import numpy as np
import torch
import torch.nn as nn

from .constraint import Constraint


class AggregateLoss(nn.Module):
    def __init__(self, objectives, constraints):
        super().__init__()
        for c in constraints:
            assert isinstance(c, Constraint)
        self.objectives = nn.ModuleList(objectives)
        self.constraints = nn.ModuleList(constraints)
        keys = []
        for term in [*self.objectives, *self.constraints]:
            keys += term.input_keys
        self.input_keys = list(set(keys))
        self.output_keys = ["loss", "objective_loss", "penalty_loss", "C_violations", "C_values",
                            "C_eq_violations", "C_ineq_violations", "C_eq_values", "C_ineq_values"]
        self._check_keys()

    def _check_keys(self):
        seen = set()
        for term in [*self.objectives, *self.constraints]:
            seen |= set(term.input_keys)
            clash = set(term.output_keys) & seen
            assert not clash, f"Keys {clash} are being overwritten by the loss term {term}."
            seen |= set(term.output_keys)

    def terms(self):
        """All terms in the order their scalars are summed: objectives, then constraints."""
        return [*self.objectives, *self.constraints]

    # -- eager evaluation -------------------------------------------------------------------
    def calculate_objectives(self, input_dict):
        out, total = {}, 0.0
        for obj in self.objectives:
            res = obj(input_dict)
            if isinstance(res, torch.Tensor):
                res = {obj.output_keys[0]: res}
            out.update(res)
            total = total + out[obj.output_keys[0]]
        out["objective_loss"] = total
        return out

    @staticmethod
    def constraint_matrices(constraints, per_term):
        """Build ``C_values`` etc. from ``[(value, violation), ...]`` (reference loss.py:86-106)."""
        vals, viols, is_eq = [], [], []
        for con, (value, violation) in zip(constraints, per_term):
            width = math.prod(value.shape[1:])
            is_eq += width * [str(con.comparator) == "eq"]
            vals.append(value.reshape(value.shape[0], -1))
            viols.append(violation.reshape(violation.shape[0], -1))
        flags = np.array(is_eq, dtype=bool)
        c_viol, c_val = torch.cat(viols, dim=-1), torch.cat(vals, dim=-1)
        return {"C_violations": c_viol, "C_values": c_val,
                "C_eq_violations": c_viol[:, flags], "C_ineq_violations": c_viol[:, ~flags],
                "C_eq_values": c_val[:, flags], "C_ineq_values": c_val[:, ~flags]}

    def calculate_constraints(self, input_dict):
        out, total, per_term = {}, 0.0, []
        for con in self.constraints:
            res = con(input_dict)
            out.update(res)
            total = total + res[con.output_keys[0]]
            per_term.append((res[con.output_keys[1]], res[con.output_keys[2]]))
        if len(self.constraints):
            out.update(self.constraint_matrices(self.constraints, per_term))
        out["penalty_loss"] = total
        return out

    def forward(self, input_dict):  # pragma: no cover - abstract
        raise NotImplementedError

    # -- algebra ----------------------------------------------------------------------------
    def __add__(self, other):
        if not isinstance(other, type(self)):
            raise ValueError("Only instances of the same loss can be added")
        import copy

        def tagged(terms, owner):
            dup = [copy.copy(t) for t in terms]
            for t in dup:
                t.name += f"_{id(owner)}"
            return dup
        return type(self)(tagged(self.objectives, self) + tagged(other.objectives, other),
                          tagged(self.constraints, self) + tagged(other.constraints, other))

    def __mul__(self, weight):
        return type(self)([weight * o for o in self.objectives], [weight * c for c in self.constraints])
    __rmul__ = __mul__


class PenaltyLoss(AggregateLoss):
    """Penalty method: ``loss = objective_loss + penalty_loss`` (reference loss.py:168-181)."""

    def forward(self, input_dict):
        objs = self.calculate_objectives(input_dict)
        merged = {**input_dict, **objs}
        pens = self.calculate_constraints(merged)
        merged.update(pens)
        merged["loss"] = objs["objective_loss"] + pens["penalty_loss"]
        return merged


class BarrierLoss(PenaltyLoss):
    """Barrier method (reference loss.py:184-257): a satisfied constraint (value < 0) contributes
    ``weight * mean(barrier(value))``, clamped to ``[0, upper_bound]``, a violated one its penalty.

    With one of the named barrier functions the constraints are lowered to ``NmTerm`` descriptors carrying the barrier
    kind (``NM_BARRIER_*``) and scored / differentiated inside the fused kernels like PenaltyLoss terms
    (``plan.py``; ``penalty_loss`` and ``loss`` come from the kernels, the per-constraint plain penalties, the
    ``<name>_barrier`` tensors and ``barrier_loss`` are materialised lazily).  With a callable barrier, or a constraint
    that cannot be lowered, this loss is evaluated with torch ops on the CUDA trajectories the fused rollout wrote and
    its gradient enters the fused adjoint kernel through ``grad_*_ext`` (``forward`` below)."""

    def __init__(self, objectives, constraints, barrier="log10", upper_bound=1., shift=1., alpha=0.5):
        super().__init__(objectives, constraints)
        self.shift, self.alpha = shift, alpha
        self.barriers = {"log10": lambda value: -torch.log10(-value),
                         "log": lambda value: -torch.log(-value),
                         "inverse": lambda value: 1 / (-value),
                         "softexp": lambda value: (torch.exp(self.alpha * value) - 1) / self.alpha + self.alpha,
                         "softlog": lambda value: -torch.log(1 + self.alpha * (-value - self.alpha)) / self.alpha,
                         "expshift": lambda value: torch.exp(value + self.shift)}
        self.barrier_name = barrier if isinstance(barrier, str) and barrier in self.barriers else None     # (None: a callable)
        if barrier in self.barriers:
            self.barrier = self.barriers[barrier]
        else:
            assert callable(barrier), f"The barrier, {barrier} must be a key in {self.barriers} or a callable."
            self.barrier = barrier
        self.upper_bound = upper_bound

    def calculate_constraints(self, input_dict):
        loss, b_loss = 0.0, 0.0
        out = super().calculate_constraints(input_dict)
        for c in self.constraints:
            cvalue, cviolation = out[c.output_keys[1]], out[c.output_keys[2]]
            penalty_mask = cvalue >= 0
            cbarrier = self.barrier(cvalue)
            cbarrier[cbarrier != cbarrier] = 0.0               # nan -> 0: infeasible entries
            cbarrier[cbarrier == float("Inf")] = 0.0           # inf -> 0: active constraints
            cbarrier = torch.clamp(cbarrier, min=0.0, max=self.upper_bound)
            out[f"{c.name}_barrier"] = cbarrier
            if penalty_mask.any():
                loss = loss + c.weight * torch.mean(penalty_mask * cviolation)
            if (~penalty_mask).any():
                barrier_loss = c.weight * torch.mean(~penalty_mask * cbarrier)
                b_loss = b_loss + barrier_loss
                loss = loss + barrier_loss
        out["barrier_loss"] = b_loss
        out["penalty_loss"] = loss
        return out
