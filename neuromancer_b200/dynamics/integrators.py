"""
Fixed-step explicit integrators on the hot path: ``Euler``, ``RK2``, ``RK4`` and ``Euler_Trap``,
``RK4_Trap``, ``Luther``, ``Runge_Kutta_Fehlberg``, and the second-order schemes ``LeapFrog`` / ``Yoshida4``.

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

NM_INT_EULER, NM_INT_RK4, NM_INT_RK2, NM_INT_EULER_TRAP, NM_INT_RK4_TRAP, NM_INT_LUTHER, NM_INT_RKF = 0, 1, 2, 3, 4, 5, 6
NM_INT_LEAPFROG, NM_INT_YOSHIDA4 = 7, 8


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


class Euler_Trap(Integrator):
    """Forward-Euler predictor, trapezoidal corrector (reference integrators.py:159-177)."""
    nm_integrator_id = NM_INT_EULER_TRAP

    def integrate(self, x, *args):
        f, h = self.block, self.h
        slope = f(x, *args)
        return x + 0.5 * h * (slope + f(x + h * slope, *args))


class RK4_Trap(Integrator):
    """RK4 predictor, trapezoidal corrector (reference integrators.py:214-235)."""
    nm_integrator_id = NM_INT_RK4_TRAP

    def integrate(self, x, *args):
        f, h = self.block, self.h
        s1 = f(x, *args)
        s2 = f(x + h * s1 / 2.0, *args)
        s3 = f(x + h * s2 / 2.0, *args)
        s4 = f(x + h * s3, *args)
        pred = x + h * (s1 / 6.0 + s2 / 3.0 + s3 / 3.0 + s4 / 6.0)
        return x + 0.5 * h * (s1 + f(pred, *args))


def _butcher_step(f, x, h, table, weights, args):
    """Generic explicit Runge-Kutta step from a strictly lower-triangular tableau."""
    ks = []
    for row in table:
        xs = x
        if row:
            xs = x + h * sum(a * k for a, k in zip(row, ks) if a != 0)
        ks.append(f(xs, *args))
    return x + h * sum(b * k for b, k in zip(weights, ks) if b != 0)


class Luther(Integrator):
    """Luther's 7-stage 6th-order method (reference integrators.py:238-264)."""
    nm_integrator_id = NM_INT_LUTHER

    def integrate(self, x, *args):
        q = 21 ** 0.5
        table = [[], [1.0], [3 / 8, 1 / 8], [8 / 27, 2 / 27, 8 / 27],
                 [(-21 + 9 * q) / 392, (-56 + 8 * q) / 392, (336 - 48 * q) / 392, (-63 + 3 * q) / 392],
                 [(-1155 - 255 * q) / 1960, (-280 - 40 * q) / 1960, -320 * q / 1960, (63 + 363 * q) / 1960, (2352 + 392 * q) / 1960],
                 [(330 + 105 * q) / 180, 120 / 180, (-200 + 280 * q) / 180, (126 - 189 * q) / 180, (-686 - 126 * q) / 180, (490 - 70 * q) / 180]]
        return _butcher_step(self.block, x, self.h, table, [1 / 20, 0, 16 / 45, 0, 49 / 180, 49 / 180, 1 / 20], args)


class Runge_Kutta_Fehlberg(Integrator):
    """Runge-Kutta-Fehlberg 4(5); returns the 5th-order solution (reference integrators.py:267-301).
    The reference also appends the 5th-4th order difference to ``local_error``; the fused path does not
    materialise that diagnostic."""
    nm_integrator_id = NM_INT_RKF

    def __init__(self, block, interp_u=None, h=1.0):
        super().__init__(block=block, interp_u=interp_u, h=h)
        self.local_error = []

    def integrate(self, x, *args):
        table = [[], [1 / 4], [3 / 32, 9 / 32], [1932 / 2197, -7200 / 2197, 7296 / 2197],
                 [439 / 216, -8.0, 3680 / 513, -845 / 4104], [-8 / 27, 2.0, -3544 / 2565, 1859 / 4104, -11 / 40]]
        return _butcher_step(self.block, x, self.h, table, [16 / 135, 0, 6656 / 12825, 28561 / 56430, -9 / 50, 2 / 55], args)


class LeapFrog(Integrator):
    """Leapfrog / velocity-Verlet step for ``ddx = f(x)`` on the state ``X = [x, dx]`` (reference integrators.py:340-360).
    ``block`` maps the ``SysDim`` positions (plus inputs) to ``SysDim`` accelerations."""
    nm_integrator_id = NM_INT_LEAPFROG
    nm_second_order = True

    def integrate(self, X, *args):
        import torch
        n = X.shape[-1] // 2
        x, dx = X[:, :n], X[:, n:2 * n]
        acc = self.block(x, *args)
        x_new = x + dx * self.h + 0.5 * acc * self.h ** 2
        dx_new = dx + 0.5 * (acc + self.block(x_new, *args)) * self.h
        return torch.cat([x_new, dx_new], dim=-1)


class Yoshida4(Integrator):
    """4th-order Yoshida composition for ``ddx = f(x)`` on ``X = [x, dx]`` (reference integrators.py:363-404)."""
    nm_integrator_id = NM_INT_YOSHIDA4
    nm_second_order = True

    def integrate(self, X, *args):
        import torch
        n = X.shape[-1] // 2
        q, v = X[:, :n], X[:, n:2 * n]
        w0 = -2 ** (1 / 3) / (2 - 2 ** (1 / 3))
        w1 = 1 / (2 - 2 ** (1 / 3))
        drift = [w1 / 2, (w0 + w1) / 2, (w0 + w1) / 2, w1 / 2]
        kick = [w1, w0, w1]
        for c, d in zip(drift[:3], kick):
            q = q + c * v * self.h
            v = v + d * self.block(q, *args) * self.h
        q = q + drift[3] * v * self.h
        return torch.cat([q, v], dim=-1)


integrators = {"LeapFrog": LeapFrog, "Yoshida4": Yoshida4, "Euler": Euler, "Euler_Trap": Euler_Trap, "RK2": RK2, "RK4": RK4, "RK4_Trap": RK4_Trap,
               "Luther": Luther, "Runge_Kutta_Fehlberg": Runge_Kutta_Fehlberg}
