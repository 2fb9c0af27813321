"""
Fixed-step explicit integrators on the hot path: ``Euler``, ``RK2``, ``RK4``.

API mirror of the reference's ``Integrator`` family (``dynamics/integrators.py``: base :14-43,
``Euler`` :144-156, ``RK2`` :180-193, ``RK4`` :196-211).  Inside a ``System`` an integrator node is
lowered to ``(NM_INT_*, h, NM_RHS_*)`` and the stages are evaluated in registers by the fused
kernels with the control input held constant over the step (zero-order hold), in exactly the
reference's operation order:

    RK4:  k1 = f(x); k2 = f(x + h*k1/2); k3 = f(x + h*k2/2); k4 = f(x + h*k3)
          x+ = x + h*(k1/6 + k2/3 + k3/3 + k4/6)

``integrate`` below is the eager single-step form for calls outside a rollout.
"""
import warnings

import torch.nn as nn

NM_INT_EULER, NM_INT_RK4, NM_INT_RK2 = 0, 1, 2


class Integrator(nn.Module):
    #: kernel id, set by subclasses that the fused path implements
    nm_integrator_id = None

    def __init__(self, block, interp_u=None, h=1.0):
        super().__init__()
        self.block = block
        self.in_features, self.out_features = block.in_features, block.out_features
        self.h = h
        if interp_u is not None:
            warnings.warn("interp_u is deprecated and has no effect.", FutureWarning)

    def step_size(self):
        """``h`` as a python float (the reference allows a 0-dim tensor, e.g.
        ``RK4(ode, h=torch.tensor(ts))``)."""
        return float(self.h)

    def integrate(self, x, *args):  # pragma: no cover - abstract
        raise NotImplementedError

    def forward(self, x, *args):
        return self.integrate(x, *args)

    def reg_error(self):
        return sum(c.reg_error() for c in self.children() if hasattr(c, "reg_error"))


class Euler(Integrator):
    nm_integrator_id = NM_INT_EULER

    def integrate(self, x, *args):
        return x + self.h * self.block(x, *args)


class RK2(Integrator):
    nm_integrator_id = NM_INT_RK2

    def integrate(self, x, *args):
        f, h = self.block, self.h
        mid = x + h * f(x, *args) / 2.0
        return x + h * f(mid, *args)


class RK4(Integrator):
    nm_integrator_id = NM_INT_RK4

    def integrate(self, x, *args):
        f, h = self.block, self.h
        s1 = f(x, *args)
        s2 = f(x + h * s1 / 2.0, *args)
        s3 = f(x + h * s2 / 2.0, *args)
        s4 = f(x + h * s3, *args)
        return x + h * (s1 / 6.0 + s2 / 3.0 + s3 / 3.0 + s4 / 6.0)


integrators = {"Euler": Euler, "RK2": RK2, "RK4": RK4}
