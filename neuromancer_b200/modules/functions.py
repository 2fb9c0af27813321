"""
Output-bounding maps used by ``MLP_bounds``.

Mirrors the two bound methods of the reference (``modules/functions.py:9-34``).  Inside a fused
rollout they are lowered to ``NM_BOUNDS_SIGMOID_SCALE`` / ``NM_BOUNDS_RELU_CLAMP`` and evaluated by
the CUDA kernels; the eager forms below serve single-module calls outside a ``System``.
"""
import torch


def bounds_scaling(x, xmin, xmax, scaling=1.0):
    """Squash ``x`` into ``[xmin, xmax]`` with a scaled logistic (reference functions.py:9-19)."""
    span = xmax - xmin
    return span * torch.sigmoid(scaling * x) + xmin


def bounds_clamp(x, xmin=None, xmax=None):
    """ReLU-based clamp, differentiable at the kinks the way the reference's is
    (functions.py:22-34): the lower bound is applied first, then the upper bound."""
    if xmin is not None:
        x = x + torch.relu(xmin - x)
    if xmax is not None:
        x = x - torch.relu(x - xmax)
    return x


functions = {"bounds_scaling": bounds_scaling, "bounds_clamp": bounds_clamp}
