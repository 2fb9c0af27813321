"""
State-space models and ODE right-hand sides on the hot path.

API mirror of the reference (``dynamics/ode.py``: ``SSM`` :7-51, ``ODESystem`` :54-72,
``TwoTankParam`` :216-237, ``VanDerPolControl`` :352-368).  Each class carries the kernel id
(``NM_RHS_*``) and the constants the plan compiler copies into ``NmPlan.rhs_const``; the fused
kernels evaluate the equations in registers.  ``ode_equations`` is the eager form for
single-module calls.

``CoupledVanDerPolControl`` is this repository's synthetic 12-state plant of benchmark config C5
(SURVEY.md section 8(d)): a ring of ``nx/2`` ``VanDerPolControl`` oscillators with nearest-neighbour
linear coupling ``kappa * (x1[j+1] - x1[j])``.
"""
import torch
import torch.nn as nn

from ..slim.linear import DerivedWeight

NM_RHS_TWOTANK, NM_RHS_VDP, NM_RHS_CVDP, NM_RHS_MLP, NM_RHS_LINEAR = 1, 2, 3, 4, 5
NM_RHS_LORENZ_CTRL, NM_RHS_CSTR, NM_RHS_DUFFING, NM_RHS_BRUSSELATOR, NM_RHS_LOTKA_VOLTERRA, NM_RHS_LORENZ = 6, 7, 8, 9, 10, 11
NM_RHS_BRUSS_HYBRID, NM_RHS_LV_HYBRID = 12, 13


class SSM(nn.Module):
    """Discrete-time ``x+ = fx(x) + fu(u) [+ fd(d)]`` (reference ode.py:7-51).  It is lowered to
    ``NM_DYN_SSM`` when ``fx``/``fu``/``fd`` are bias-free dense linear maps."""

    def __init__(self, fx, fu, nx, nu, fd=None, nd=0):
        super().__init__()
        self.fx, self.fu, self.fd = fx, fu, fd
        self.nx, self.nu, self.nd = nx, nu, nd
        self.in_features, self.out_features = nx + nu + nd, nx

    def forward(self, x, u, d=None):
        assert x.dim() == 2 and u.dim() == 2
        nxt = self.fx(x) + self.fu(u)
        if self.fd is not None and d is not None:
            assert d.dim() == 2
            nxt = nxt + self.fd(d)
        return nxt


class ODESystem(nn.Module):
    """Base class of continuous-time right-hand sides ``dx/dt = f(x, u)``."""

    nm_rhs_id = None

    def __init__(self, insize, outsize):
        super().__init__()
        self.in_features, self.out_features = insize, outsize
        self.nx = outsize
        self.nu = insize - outsize

    def ode_equations(self, x, *args):  # pragma: no cover - abstract
        raise NotImplementedError

    def rhs_constants_tensor(self):
        """The constants handed to the kernel (``NmPlan.rhs_const``) as ONE 1-d tensor, formed differentiably from the
        module's parameters in the reference's own operation order: the kernels' gradient w.r.t. it (system identification)
        flows on through this expression to the parameters."""
        return torch.zeros(0)

    def rhs_constants(self):
        """``rhs_constants_tensor()`` as python floats."""
        return [float(v) for v in self.rhs_constants_tensor().detach().cpu().reshape(-1)]

    def forward(self, x, *args):
        assert x.dim() == 2
        return self.ode_equations(x, *args)


class TwoTankParam(ODESystem):
    """Two-tank level dynamics with pump and valve inputs (reference ode.py:216-237)."""

    nm_rhs_id = NM_RHS_TWOTANK

    def __init__(self, insize=4, outsize=2):
        super().__init__(insize=insize, outsize=outsize)
        self.c1 = nn.Parameter(torch.tensor([0.1]), requires_grad=True)
        self.c2 = nn.Parameter(torch.tensor([0.1]), requires_grad=True)

    def rhs_constants_tensor(self):
        return torch.cat([self.c1.reshape(1), self.c2.reshape(1)])

    def ode_equations(self, x, u):
        lvl = torch.clip(x, min=0, max=1.0)
        act = torch.clip(u, min=0, max=1.0)
        h1, h2 = lvl[:, 0:1], lvl[:, 1:2]
        pump, valve = act[:, 0:1], act[:, 1:2]
        out1 = self.c2 * torch.sqrt(h1)
        d1 = self.c1 * (1.0 - valve) * pump - out1
        d2 = self.c1 * valve * pump + out1 - self.c2 * torch.sqrt(h2)
        return torch.cat([d1, d2], dim=-1)


class VanDerPolControl(ODESystem):
    """Controlled Van der Pol oscillator (reference ode.py:352-368)."""

    nm_rhs_id = NM_RHS_VDP

    def __init__(self, insize=3, outsize=2):
        super().__init__(insize=insize, outsize=outsize)
        self.mu = nn.Parameter(torch.tensor([1.0]), requires_grad=True)

    def rhs_constants_tensor(self):
        return torch.cat([self.mu.reshape(1)])

    def ode_equations(self, x, u):
        pos, vel = x[:, 0:1], x[:, 1:2]
        return torch.cat([vel, self.mu * (1 - pos ** 2) * vel - pos + u], dim=-1)


class CoupledVanDerPolControl(ODESystem):
    """Ring of ``n_osc`` controlled Van der Pol oscillators, state ordered
    ``[x1_0, x2_0, x1_1, x2_1, ...]``, one input per oscillator:

        dx1_j = x2_j
        dx2_j = mu*(1 - x1_j**2)*x2_j - x1_j + u_j + kappa*(x1_{(j+1) % n} - x1_j)
    """

    nm_rhs_id = NM_RHS_CVDP

    def __init__(self, n_osc=6, mu=1.0, kappa=0.1):
        super().__init__(insize=3 * n_osc, outsize=2 * n_osc)
        self.n_osc = n_osc
        self.mu = nn.Parameter(torch.tensor([float(mu)]), requires_grad=False)
        self.kappa = nn.Parameter(torch.tensor([float(kappa)]), requires_grad=False)

    def rhs_constants_tensor(self):
        return torch.cat([self.mu.reshape(1), self.kappa.reshape(1)])

    def ode_equations(self, x, u):
        pos, vel = x[:, 0::2], x[:, 1::2]
        nxt = torch.roll(pos, shifts=-1, dims=1)
        acc = self.mu * (1 - pos ** 2) * vel - pos + u + self.kappa * (nxt - pos)
        return torch.stack([vel, acc], dim=-1).reshape(x.shape[0], -1)


class LorenzControl(ODESystem):
    """Controlled Lorenz system, two inputs (reference ode.py:393-414)."""
    nm_rhs_id = NM_RHS_LORENZ_CTRL

    def __init__(self, insize=5, outsize=3):
        super().__init__(insize=insize, outsize=outsize)
        self.rho = nn.Parameter(torch.tensor([28.0]), requires_grad=True)
        self.sigma = nn.Parameter(torch.tensor([10.0]), requires_grad=True)
        self.beta = nn.Parameter(torch.tensor([2.66667]), requires_grad=True)

    def rhs_constants_tensor(self):
        return torch.cat([self.rho.reshape(1), self.sigma.reshape(1), self.beta.reshape(1)])

    def ode_equations(self, x, u):
        a, b, c = x[:, 0:1], x[:, 1:2], x[:, 2:3]
        return torch.cat([self.sigma * (b - a) + u[:, 0:1], a * (self.rho - c) - b, a * b - self.beta * c - u[:, 1:2]], dim=-1)


class LorenzParam(ODESystem):
    """Autonomous Lorenz system (reference ode.py:371-391)."""
    nm_rhs_id = NM_RHS_LORENZ

    def __init__(self, insize=3, outsize=3):
        super().__init__(insize=insize, outsize=outsize)
        self.rho = nn.Parameter(torch.tensor([5.0]), requires_grad=True)
        self.sigma = nn.Parameter(torch.tensor([5.0]), requires_grad=True)
        self.beta = nn.Parameter(torch.tensor([5.0]), requires_grad=True)

    def rhs_constants_tensor(self):
        return torch.cat([self.rho.reshape(1), self.sigma.reshape(1), self.beta.reshape(1)])

    def ode_equations(self, x):
        a, b, c = x[:, 0:1], x[:, 1:2], x[:, 2:3]
        return torch.cat([self.sigma * (b - a), a * (self.rho - c) - b, a * b - self.beta * c], dim=-1)


class CSTR_Param(ODESystem):
    """Exothermic CSTR with cooling-jacket temperature as input (reference ode.py:417-461)."""
    nm_rhs_id = NM_RHS_CSTR

    def __init__(self, insize=3, outsize=2):
        super().__init__(insize=insize, outsize=outsize)
        mk = lambda v, g: nn.Parameter(torch.tensor([v]), requires_grad=g)
        # the reference writes nn.Parameter(torch.tensor([v], requires_grad=<flag>)): the inner flag is dropped, every
        # constant of this class is a trainable Parameter there -- mirrored as it behaves
        self.q, self.V, self.rho, self.Cp = mk(100.0, True), mk(100.0, True), mk(1000.0, True), mk(0.239, True)
        self.mdelH, self.EoverR, self.k0 = mk(5e4, True), mk(8750.0, True), mk(7.2e10, True)
        self.UA, self.Tf, self.Caf = mk(5e4, True), mk(350.0, True), mk(1.0, True)

    def rhs_constants_tensor(self):
        # the coefficient groups are formed in fp32 exactly as the reference forms them per call
        qV = self.q / self.V
        cH = self.mdelH / (self.rho * self.Cp)
        cU = self.UA / self.V / self.rho / self.Cp
        return torch.cat([t.reshape(1) for t in (qV, self.Caf, self.Tf, self.k0, self.EoverR, cH, cU)])

    def ode_equations(self, x, u):
        conc, temp = x[:, 0:1], x[:, 1:2]
        rate = self.k0 * torch.exp(-self.EoverR / temp) * conc
        d_conc = self.q / self.V * (self.Caf - conc) - rate
        d_temp = (self.q / self.V * (self.Tf - temp) + self.mdelH / (self.rho * self.Cp) * rate
                  + self.UA / self.V / self.rho / self.Cp * (u - temp))
        return torch.cat([d_conc, d_temp], dim=-1)


class DuffingParam(ODESystem):
    """Forced Duffing oscillator; the input is time (reference ode.py:240-262)."""
    nm_rhs_id = NM_RHS_DUFFING

    def __init__(self, insize=3, outsize=2):
        super().__init__(insize=insize, outsize=outsize)
        mk = lambda v, g: nn.Parameter(torch.tensor([v]), requires_grad=g)
        self.alpha, self.beta, self.delta, self.gamma, self.omega = mk(1.0, False), mk(5.0, False), mk(0.02, False), mk(8.0, False), mk(0.5, True)

    def rhs_constants_tensor(self):
        return torch.cat([self.alpha.reshape(1), self.beta.reshape(1), self.delta.reshape(1), self.gamma.reshape(1), self.omega.reshape(1)])

    def ode_equations(self, x, u):
        pos, vel = x[:, 0:1], x[:, 1:2]
        return torch.cat([vel, -self.delta * vel - self.alpha * pos - self.beta * pos ** 3 + self.gamma * torch.cos(self.omega * u)], dim=-1)


class BrusselatorParam(ODESystem):
    """Autonomous Brusselator (reference ode.py:265-281)."""
    nm_rhs_id = NM_RHS_BRUSSELATOR

    def __init__(self, insize=2, outsize=2):
        super().__init__(insize=insize, outsize=outsize)
        self.alpha = nn.Parameter(torch.tensor([5.0]), requires_grad=True)
        self.beta = nn.Parameter(torch.tensor([5.0]), requires_grad=True)

    def rhs_constants_tensor(self):
        return torch.cat([self.alpha.reshape(1), self.beta.reshape(1)])

    def ode_equations(self, x):
        a, b = x[:, 0:1], x[:, 1:2]
        return torch.cat([self.alpha + b * a ** 2 - self.beta * a - a, self.beta * a - b * a ** 2], dim=-1)


class LotkaVolterraParam(ODESystem):
    """Autonomous predator-prey model (reference ode.py:333-349)."""
    nm_rhs_id = NM_RHS_LOTKA_VOLTERRA

    def __init__(self, insize=2, outsize=2):
        super().__init__(insize=insize, outsize=outsize)
        self.alpha = nn.Parameter(torch.tensor([1.0]), requires_grad=True)

    def rhs_constants_tensor(self):
        return torch.cat([self.alpha.reshape(1)])

    def ode_equations(self, x):
        a, b = x[:, 0:1], x[:, 1:2]
        return torch.cat([self.alpha * a - 0.4 * a * b, 0.1 * a * b - 0.4 * b], dim=-1)


class HybridODESystem(ODESystem):
    """Grey-box right-hand sides: white-box equations with one unknown term ``block(x)`` (an MLP 2 -> 1) learned from data
    next to the constants (reference ode.py:284-331)."""

    def __init__(self, block, insize=2, outsize=2):
        super().__init__(insize=insize, outsize=outsize)
        self.block = block
        assert self.block.in_features == 2
        assert self.block.out_features == 1


class BrusselatorHybrid(HybridODESystem):
    """Brusselator with the ``x2 * x1**2`` term replaced by ``block(x)`` (reference ode.py:284-304)."""
    nm_rhs_id = NM_RHS_BRUSS_HYBRID

    def __init__(self, block, insize=2, outsize=2):
        super().__init__(block, insize=insize, outsize=outsize)
        self.alpha = nn.Parameter(torch.tensor([5.0]), requires_grad=True)
        self.beta = nn.Parameter(torch.tensor([5.0]), requires_grad=True)

    def rhs_constants_tensor(self):
        return torch.cat([self.alpha.reshape(1), self.beta.reshape(1)])

    def ode_equations(self, x):
        a = x[:, 0:1]
        m = self.block(x)
        return torch.cat([self.alpha + m - self.beta * a - a, self.beta * a - m], dim=-1)


class LotkaVolterraHybrid(HybridODESystem):
    """Predator-prey model with the interaction term ``x1 * x2`` replaced by ``block(x)`` (reference ode.py:307-331)."""
    nm_rhs_id = NM_RHS_LV_HYBRID

    def __init__(self, block, insize=2, outsize=2):
        super().__init__(block, insize=insize, outsize=outsize)
        mk = lambda: nn.Parameter(torch.tensor([0.10]), requires_grad=True)
        self.alpha, self.beta, self.delta, self.gamma = mk(), mk(), mk(), mk()

    def rhs_constants_tensor(self):
        return torch.cat([t.reshape(1) for t in (self.alpha, self.beta, self.delta, self.gamma)])

    def ode_equations(self, x):
        a, b = x[:, 0:1], x[:, 1:2]
        m = self.block(x)
        return torch.cat([self.alpha * a - self.beta * m, self.delta * m - self.gamma * b], dim=-1)


class DerivedConstants(DerivedWeight):
    """The constant vector of a white-box right-hand side with trainable parameters as the fused path sees it (system
    identification; reference ode.py declares c1, c2, mu, rho ... ``nn.Parameter(requires_grad=True)``): a function of the
    module's parameters re-evaluated differentiably on every rollout call.  The kernels read the values from
    ``NmPlan.rhs_const`` (refreshed by the call) and write d loss / d constants into this entry's slot of the flat
    gradient; autograd carries it on to the parameters."""

    def __init__(self, module):
        self.module = module
        self.shape = torch.Size((len(module.rhs_constants()),))

    def numel(self):
        return self.shape[0]

    def own_parameters(self):
        """the plant's constants (a hybrid right-hand side's ``block`` is lowered as an MLP of its own)"""
        return [t for n, t in self.module.named_parameters() if not n.startswith("block.")]

    def materialise(self):
        return self.module.rhs_constants_tensor().to(torch.float32)


ode = {"TwoTankParam": TwoTankParam, "LorenzControl": LorenzControl, "LorenzParam": LorenzParam, "CSTR_Param": CSTR_Param,
       "DuffingParam": DuffingParam, "BrusselatorParam": BrusselatorParam, "LotkaVolterraParam": LotkaVolterraParam, "VanDerPolControl": VanDerPolControl,
       "CoupledVanDerPolControl": CoupledVanDerPolControl, "BrusselatorHybrid": BrusselatorHybrid, "LotkaVolterraHybrid": LotkaVolterraHybrid}
