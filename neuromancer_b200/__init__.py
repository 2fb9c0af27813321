"""
neuromancer_b200 -- B200-native (sm_100a) fused DPC rollout behind the pnnl/neuromancer API.

One hot path is implemented: the batched closed-loop horizon rollout of
``System.forward`` -- MLP policy ``Node`` -> ``RK4`` / ``Euler`` / linear SSM dynamics -> ``PenaltyLoss``
scoring -- and its backward-through-time pass, as two hand-written CUDA kernels reached through a
C-ABI shared library (``include/nm_rollout.h``).  The classes below keep the reference's names and
signatures so a DPC / system-ID script switches by changing its imports.
"""
from . import constraint, dataset, dynamics, loss, modules, problem, slim, system  # noqa: F401
from .constraint import Constraint, Objective, Variable, variable  # noqa: F401
from .dataset import DictDataset  # noqa: F401
from .loss import BarrierLoss, PenaltyLoss  # noqa: F401
from .problem import Problem  # noqa: F401
from .system import Node, System, SystemPreview  # noqa: F401

__version__ = "0.1.0"
