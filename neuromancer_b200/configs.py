"""
The benchmark problems of BASELINE.json (C1..C5) and a few small shape classes used by the tests,
written against this package's public API exactly the way the reference examples write them:

  c1  examples/control/Part_1_stabilize_linear_system.py:14-67   double integrator, nsteps 50
  c2  examples/control/Part_3_ref_tracking_ODE.py:32-163         TwoTank + RK4 reference tracking
  c3  examples/ODEs/Part_1_NODE.py:58-104                        neural ODE system ID (RK4 over an MLP RHS)
  c4  examples/domain_examples/grid_response.ipynb cells 7-30    building thermal linear SSM, 48-step preview
  c5  examples/control/Part_2_stabilize_ODE.py:33-135 scaled up  12-state coupled Van der Pol, 4x128 policy

Inputs are synthetic: shapes, value ranges and constants follow the
examples; there is no network for the building / price data of c4, so its matrices are a fixed
Schur-stable synthetic system with the example's dimensions.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict

import torch
import torch.nn as nn

from .constraint import variable
from .dynamics import integrators, ode
from .loss import PenaltyLoss
from .modules import blocks
from .modules.activations import activations
from .problem import Problem
from .system import Node, System, SystemPreview


@dataclass
class Case:
    name: str
    problem: Problem
    system: System
    nsteps: int
    default_batch: int
    make_batch: Callable[[int, object, int], Dict[str, torch.Tensor]]
    flop_per_traj_step: float      # fwd+bwd algorithmic FLOPs
    bytes_per_traj_step: float

    def plan(self, batch=None):
        if batch is None:
            batch = self.make_batch(4, "cpu", 0)
        return self.system.plan_for(self.system.init(dict(batch)), loss=self.problem.loss)

    def policy_parameters(self):
        return [p for p in self.problem.parameters() if p.requires_grad]


def _frozen_linear(weight: torch.Tensor) -> nn.Linear:
    lin = nn.Linear(weight.shape[1], weight.shape[0], bias=False)
    with torch.no_grad():
        lin.weight.copy_(weight)
    lin.weight.requires_grad_(False)
    return lin


def _gen(device, seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


# ------------------------------------------------------------------------------------------- C1
def c1(seed=0) -> Case:
    torch.manual_seed(seed)
    nx, nu, nsteps = 2, 1, 50
    A = torch.tensor([[1.2, 1.0], [0.0, 1.0]])
    B = torch.tensor([[1.0], [0.5]])
    mlp = blocks.MLP(nx, nu, bias=True, linear_map=nn.Linear, nonlin=nn.ReLU, hsizes=[20, 20, 20, 20])
    policy = Node(mlp, ["X"], ["U"], name="policy")
    # the example's lambda  x @ A.T + u @ B.T  expressed as a recognisable ode.SSM
    ssm = ode.SSM(_frozen_linear(A), _frozen_linear(B), nx, nu)
    plant = Node(ssm, ["X", "U"], ["X"], name="double_integrator")
    system = System([policy, plant], nsteps=nsteps, name="cl_system")
    u, x = variable("U"), variable("X")
    action_loss = 0.0001 * (u == 0.) ^ 2
    regulation_loss = 10. * (x == 0.) ^ 2
    action_loss.name, regulation_loss.name = "action_loss", "regulation_loss"
    problem = Problem([system], PenaltyLoss([action_loss, regulation_loss], []))

    def make_batch(Bn, device, s=0):
        g = _gen(device, 1000 + s)
        return {"X": (3. * torch.randn(Bn, 1, nx, generator=g)).to(device), "name": "train"}
    return Case("c1", problem, system, nsteps, 3333, make_batch, 7.6e3, 24.0)


# ------------------------------------------------------------------------------------------- C2
def c2(seed=0, nsteps=30) -> Case:
    torch.manual_seed(seed)
    nx, nu, nref = 2, 2, 2
    two_tank = ode.TwoTankParam()
    two_tank.c1 = nn.Parameter(torch.tensor(0.08), requires_grad=False)     # psl TwoTank constants
    two_tank.c2 = nn.Parameter(torch.tensor(0.04), requires_grad=False)
    integrator = integrators.RK4(two_tank, h=torch.tensor(1.0))
    model = Node(integrator, ["x", "u"], ["x"], name="model")
    net = blocks.MLP_bounds(insize=nx + nref, outsize=nu, hsizes=[32, 32], nonlin=activations["gelu"],
                            min=0, max=1.)
    policy = Node(net, ["x", "r"], ["u"], name="policy")
    system = System([policy, model], nsteps=nsteps, name="cl_system")
    x, ref = variable("x"), variable("r")
    regulation_loss = 5. * ((x == ref) ^ 2)
    lo = 10. * (x > 0)
    hi = 10. * (x < 1.)
    t_lo = 10. * (x[:, [-1], :] > ref - 0.01)
    t_hi = 10. * (x[:, [-1], :] < ref + 0.01)
    for term, nm_ in ((regulation_loss, "ref_tracking"), (lo, "x_min"), (hi, "x_max"),
                      (t_lo, "x_N_min"), (t_hi, "x_N_max")):
        term.name = nm_
    problem = Problem([system], PenaltyLoss([regulation_loss], [lo, hi, t_lo, t_hi]))

    def make_batch(Bn, device, s=0):
        g = _gen(device, 2000 + s)
        x0 = torch.rand(Bn, 1, nx, generator=g)
        r = torch.rand(Bn, 1, 1, generator=g) * torch.ones(nsteps + 1, nref)
        return {"x": x0.to(device), "r": r.contiguous().to(device), "name": "train"}
    return Case("c2", problem, system, nsteps, 65536, make_batch, 7.8e3, 48.0)


# ------------------------------------------------------------------------------------------- C3
def c3(seed=0, nsteps=100) -> Case:
    torch.manual_seed(seed)
    nx = 2
    fx = blocks.MLP(nx, nx, bias=True, linear_map=nn.Linear, nonlin=nn.ReLU, hsizes=[60, 60, 60])
    fx_rk4 = integrators.RK4(fx, h=0.1)
    model = Node(fx_rk4, ["xn"], ["xn"], name="NODE")
    system = System([model], name="system")            # nsteps inferred from data['X'] (nstep_key)
    x = variable("X")
    xhat = variable("xn")[:, :-1, :]
    x_fd = (x[:, 1:, :] - x[:, :-1, :])
    xhat_fd = (xhat[:, 1:, :] - xhat[:, :-1, :])
    fd_loss = 2. * (x_fd == xhat_fd) ^ 2
    fd_loss.name = "FD_loss"
    reference_loss = (xhat == x) ^ 2
    reference_loss.name = "ref_loss"
    problem = Problem([system], PenaltyLoss([reference_loss, fd_loss], []))

    def make_batch(Bn, device, s=0):
        g = _gen(device, 3000 + s)
        X = torch.randn(Bn, nsteps, nx, generator=g)
        return {"X": X.to(device), "xn": X[:, 0:1, :].contiguous().to(device), "name": "train"}
    return Case("c3", problem, system, nsteps, 262144, make_batch, 178.6e3, 32.0)


# ------------------------------------------------------------------------------------------- C4
C4_A = [[0.9500, 0.0200, 0.0000, 0.0100],
        [0.0300, 0.9000, 0.0200, 0.0000],
        [0.0000, 0.0400, 0.9200, 0.0100],
        [0.0200, 0.0100, 0.0300, 0.9300]]
C4_B = [[0.00], [0.00], [0.05], [0.10]]
C4_E = [[0.010, 0.000, 0.002], [0.020, 0.001, 0.000], [0.005, 0.002, 0.001], [0.002, 0.001, 0.003]]
C4_C = [[0.0, 0.0, 0.0, 1.0]]


def c4(seed=0, nsteps=96, horizon=48) -> Case:
    torch.manual_seed(seed)
    nx, nu, nd, ny, nd_obs = 4, 1, 3, 1, 1
    A, Bm, E, Cm = (torch.tensor(m) for m in (C4_A, C4_B, C4_E, C4_C))
    ssm = ode.SSM(_frozen_linear(A), _frozen_linear(Bm), nx, nu, fd=_frozen_linear(E), nd=nd)
    state_model = Node(ssm, ["x", "u", "d"], ["x"], name="SSM")
    obs_model = Node(_frozen_linear(Cm), ["x"], ["y"], name="y=Cx")
    forecast = ["d_obs_lookahead", "ymin_lookahead", "ymax_lookahead", "price_lookahead"]
    insize = ny + horizon * (nd_obs + ny + ny + 1)
    net = blocks.MLP_bounds(insize=insize, outsize=nu, hsizes=[32, 32], nonlin=nn.LeakyReLU,
                            min=torch.tensor([0.0]), max=torch.tensor([5000.0]))
    policy = Node(net, ["y"] + forecast, ["u"], name="policy")
    system = System([obs_model, policy, state_model], nsteps=nsteps, name="cl_system")
    y_var, u_var = variable("y"), variable("u")
    ymin_var = variable("ymin_lookahead")[:, :-1, 0:1]
    ymax_var = variable("ymax_lookahead")[:, :-1, 0:1]
    price_var = variable("price_lookahead")[:, :-1, 0:1]
    price_loss = 1 * ((u_var * price_var) == torch.zeros_like(u_var)) ^ 2
    comfort_lower = 5e06 * (y_var > ymin_var)
    comfort_upper = 5e06 * (y_var < ymax_var)
    price_loss.name, comfort_lower.name, comfort_upper.name = "price_loss", "y_min", "y_max"
    problem = Problem([system], PenaltyLoss([price_loss], [comfort_lower, comfort_upper]))

    def lookahead(sig):                       # [B, nsteps+horizon, D] -> [B, nsteps+1, D*horizon]
        Bn, Tn, D = sig.shape
        win = sig.unfold(1, horizon, 1)       # [B, nsteps+1, D, horizon]
        return win.permute(0, 1, 3, 2).reshape(Bn, Tn - horizon + 1, horizon * D).contiguous()

    def make_batch(Bn, device, s=0):
        g = _gen(device, 4000 + s)
        Tn = nsteps + horizon
        ymin = (18.0 + 4.5 * torch.rand(Bn, 1, ny, generator=g)).repeat(1, Tn, 1)
        ymax = (22.5 + 4.5 * torch.rand(Bn, 1, ny, generator=g)).repeat(1, Tn, 1)
        dist = torch.randn(Bn, Tn, nd, generator=g)
        dist[:, :, 0] = 5.0 + 8.0 * dist[:, :, 0]
        price = 0.05 + 0.2 * torch.rand(Bn, Tn, 1, generator=g)
        x0 = 20.0 + torch.randn(Bn, 1, nx, generator=g)
        batch = {"x": x0, "d": dist[:, :nsteps, :].contiguous(),
                 "d_obs_lookahead": lookahead(dist[:, :, 0:1]), "ymin_lookahead": lookahead(ymin),
                 "ymax_lookahead": lookahead(ymax), "price_lookahead": lookahead(price)}
        out = {k: v.to(device) for k, v in batch.items()}
        out["name"] = "train"
        return out
    return Case("c4", problem, system, nsteps, 131072, make_batch, 43.7e3, 1600.0)


def c4p(seed=0, nsteps=96, horizon=48) -> Case:
    """C4 written with ``SystemPreview``: the policy previews the RAW disturbance / comfort-band / price
    signals (window = horizon) instead of reading materialised look-ahead tensors.  Same arithmetic as c4 --
    look-ahead column j*D+d at step t IS signal entry t+j -- with 1/48 of the exogenous HBM traffic."""
    torch.manual_seed(seed)
    nx, nu, nd, ny = 4, 1, 3, 1
    A, Bm, E, Cm = (torch.tensor(m) for m in (C4_A, C4_B, C4_E, C4_C))
    ssm = ode.SSM(_frozen_linear(A), _frozen_linear(Bm), nx, nu, fd=_frozen_linear(E), nd=nd)
    state_model = Node(ssm, ["x", "u", "d"], ["x"], name="SSM")
    obs_model = Node(_frozen_linear(Cm), ["x"], ["y"], name="y=Cx")
    raw = ["d_obs", "ymin", "ymax", "price"]
    net = blocks.MLP_bounds(insize=ny + horizon * 4, outsize=nu, hsizes=[32, 32], nonlin=nn.LeakyReLU,
                            min=torch.tensor([0.0]), max=torch.tensor([5000.0]))
    policy = Node(net, ["y"] + raw, ["u"], name="policy")
    system = SystemPreview([obs_model, policy, state_model], nsteps=nsteps, name="cl_system",
                           preview_keys_map={k: ["policy"] for k in raw}, preview_length={k: horizon - 1 for k in raw},
                           pad_mode="replicate")
    y_var, u_var = variable("y"), variable("u")
    price_loss = 1 * ((u_var * variable("price")[:, :nsteps, 0:1]) == torch.zeros_like(u_var)) ^ 2
    comfort_lower = 5e06 * (y_var > variable("ymin")[:, :nsteps, 0:1])
    comfort_upper = 5e06 * (y_var < variable("ymax")[:, :nsteps, 0:1])
    price_loss.name, comfort_lower.name, comfort_upper.name = "price_loss", "y_min", "y_max"
    problem = Problem([system], PenaltyLoss([price_loss], [comfort_lower, comfort_upper]))

    def make_batch(Bn, device, s=0):
        g = _gen(device, 4000 + s)                       # same stream as c4.make_batch -> same signals
        Tn = nsteps + horizon
        ymin = (18.0 + 4.5 * torch.rand(Bn, 1, ny, generator=g)).repeat(1, Tn, 1)
        ymax = (22.5 + 4.5 * torch.rand(Bn, 1, ny, generator=g)).repeat(1, Tn, 1)
        dist = torch.randn(Bn, Tn, nd, generator=g)
        dist[:, :, 0] = 5.0 + 8.0 * dist[:, :, 0]
        price = 0.05 + 0.2 * torch.rand(Bn, Tn, 1, generator=g)
        x0 = 20.0 + torch.randn(Bn, 1, nx, generator=g)
        batch = {"x": x0, "d": dist[:, :nsteps, :].contiguous(), "d_obs": dist[:, :, 0:1].contiguous(),
                 "ymin": ymin.contiguous(), "ymax": ymax.contiguous(), "price": price.contiguous()}
        out = {k: v.to(device) for k, v in batch.items()}
        out["name"] = "train"
        return out
    return Case("c4p", problem, system, nsteps, 131072, make_batch, 43.7e3, 100.0)


# ------------------------------------------------------------------------------------------- C5
def c5(seed=0, nsteps=200) -> Case:
    torch.manual_seed(seed)
    n_osc = 6
    nx, nu = 2 * n_osc, n_osc
    plant = ode.CoupledVanDerPolControl(n_osc=n_osc, mu=1.0, kappa=0.1)
    integrator = integrators.RK4(plant, h=torch.tensor(0.05))
    model = Node(integrator, ["x", "u"], ["x"], name="model")
    net = blocks.MLP_bounds(insize=2 * nx, outsize=nu, hsizes=[128, 128, 128, 128],
                            nonlin=activations["gelu"], min=-5., max=5.)
    policy = Node(net, ["x", "r"], ["u"], name="policy")
    system = System([policy, model], nsteps=nsteps, name="cl_system")
    x, ref, u = variable("x"), variable("r"), variable("u")
    regulation = 1. * ((x == ref) ^ 2)
    effort = 0.01 * ((u == 0.) ^ 2)
    lo, hi = 10. * (x > -4.), 10. * (x < 4.)
    t_lo = 10. * (x[:, [-1], :] > ref - 0.01)
    t_hi = 10. * (x[:, [-1], :] < ref + 0.01)
    for term, nm_ in ((regulation, "state_loss"), (effort, "action_loss"), (lo, "x_min"), (hi, "x_max"),
                      (t_lo, "x_N_min"), (t_hi, "x_N_max")):
        term.name = nm_
    problem = Problem([system], PenaltyLoss([regulation, effort], [lo, hi, t_lo, t_hi]))

    def make_batch(Bn, device, s=0):
        g = _gen(device, 5000 + s)
        return {"x": torch.randn(Bn, 1, nx, generator=g).to(device),
                "r": torch.zeros(Bn, nsteps + 1, nx, device=device), "name": "train"}
    return Case("c5", problem, system, nsteps, 4194304, make_batch, 319.5e3, 240.0)


# ------------------------------------------------------------------------------------------- test shapes
def t_vdp_rk4_tanh_clamp(seed=0, nsteps=12, linear_map=None, name="t_vdp_rk4_tanh_clamp") -> Case:
    """``linear_map``: a structured slim map for every policy layer (t_slim_*: SURVEY 8 f3) -- same shape class"""
    torch.manual_seed(seed)
    plant = ode.VanDerPolControl()
    plant.mu = nn.Parameter(torch.tensor(1.0), requires_grad=False)
    model = Node(integrators.RK4(plant, h=0.1), ["x", "u"], ["x"], name="model")
    extra = {} if linear_map is None else {"linear_map": linear_map}
    net = blocks.MLP_bounds(insize=4, outsize=1, hsizes=[16, 16], nonlin=nn.Tanh, min=-1., max=1.,
                            method="relu_clamp", **extra)
    system = System([Node(net, ["x", "r"], ["u"], name="policy"), model], nsteps=nsteps, name="sys")
    x, r, u = variable("x"), variable("r"), variable("u")
    obj = 3. * ((x == r) ^ 2)
    eff = 0.1 * (u == 0.)                     # norm-1 equality
    box = 2. * ((x < 1.5) ^ 2)                # norm-2 inequality
    du = 0.5 * ((u[:, 1:, :] - u[:, :-1, :]) == 0.) ^ 2
    for t_, n_ in ((obj, "obj"), (eff, "eff"), (box, "box"), (du, "du")):
        t_.name = n_
    problem = Problem([system], PenaltyLoss([obj, eff, du], [box]))

    def make_batch(Bn, device, s=0):
        g = _gen(device, 6000 + s)
        return {"x": torch.randn(Bn, 1, 2, generator=g).to(device),
                "r": (0.5 * torch.randn(Bn, nsteps + 1, 2, generator=g)).to(device), "name": "train"}
    return Case(name, problem, system, nsteps, 256, make_batch, 0, 0)


def t_slim_pf(seed=0, nsteps=12) -> Case:
    from . import slim
    return t_vdp_rk4_tanh_clamp(seed=seed, nsteps=nsteps, linear_map=slim.PerronFrobeniusLinear, name="t_slim_pf")


def t_slim_svd(seed=0, nsteps=12) -> Case:
    from . import slim
    return t_vdp_rk4_tanh_clamp(seed=seed, nsteps=nsteps, linear_map=slim.SVDLinear, name="t_slim_svd")


def _t_barrier(barrier, seed=0, nsteps=12) -> Case:
    """the tanh-clamp problem under a BarrierLoss (reference loss.py:184-257): box constraints on state and action"""
    from .loss import BarrierLoss
    case = t_vdp_rk4_tanh_clamp(seed=seed, nsteps=nsteps, name=f"t_barrier_{barrier}")
    x, r, u = variable("x"), variable("r"), variable("u")
    obj = 3. * ((x == r) ^ 2)
    hi = 2. * ((x < 1.5) ^ 2)
    lo = 1. * (x > -1.5)
    umax = 0.5 * (u < 0.8)
    for t_, n_ in ((obj, "obj"), (hi, "hi"), (lo, "lo"), (umax, "umax")):
        t_.name = n_
    case.problem = Problem([case.system], BarrierLoss([obj], [hi, lo, umax], barrier=barrier, upper_bound=2.0, shift=0.7, alpha=0.4))
    return case


def t_barrier_log10(seed=0, **kw): return _t_barrier("log10", seed, **kw)
def t_barrier_log(seed=0, **kw): return _t_barrier("log", seed, **kw)
def t_barrier_inverse(seed=0, **kw): return _t_barrier("inverse", seed, **kw)
def t_barrier_softexp(seed=0, **kw): return _t_barrier("softexp", seed, **kw)
def t_barrier_softlog(seed=0, **kw): return _t_barrier("softlog", seed, **kw)
def t_barrier_expshift(seed=0, **kw): return _t_barrier("expshift", seed, **kw)


def t_vdp_euler_elu(seed=0, nsteps=9) -> Case:
    torch.manual_seed(seed)
    plant = ode.VanDerPolControl()
    plant.mu = nn.Parameter(torch.tensor(0.7), requires_grad=False)
    model = Node(integrators.Euler(plant, h=0.05), ["x", "u"], ["x"], name="model")
    net = blocks.MLP(insize=2, outsize=1, hsizes=[8], nonlin=nn.ELU, bias=False)
    system = System([Node(net, ["x"], ["u"], name="policy"), model], nsteps=nsteps, name="sys")
    x, u = variable("x"), variable("u")
    obj = (x == 0.) ^ 2
    pen = 4. * (u > -0.2)
    obj.name, pen.name = "obj", "pen"
    problem = Problem([system], PenaltyLoss([obj], [pen]))

    def make_batch(Bn, device, s=0):
        g = _gen(device, 7000 + s)
        return {"x": torch.randn(Bn, 1, 2, generator=g).to(device), "name": "train"}
    return Case("t_vdp_euler_elu", problem, system, nsteps, 200, make_batch, 0, 0)


def t_twotank_rk2_sigmoid(seed=0, nsteps=7) -> Case:
    torch.manual_seed(seed)
    plant = ode.TwoTankParam()
    plant.c1 = nn.Parameter(torch.tensor(0.08), requires_grad=False)
    plant.c2 = nn.Parameter(torch.tensor(0.04), requires_grad=False)
    model = Node(integrators.RK2(plant, h=0.5), ["x", "u"], ["x"], name="model")
    net = blocks.MLP_bounds(insize=4, outsize=2, hsizes=[12], nonlin=nn.Sigmoid, min=0., max=1.)
    system = System([Node(net, ["x", "r"], ["u"], name="policy"), model], nsteps=nsteps, name="sys")
    x, r = variable("x"), variable("r")
    obj = (x == r) ^ 2
    obj.name = "obj"
    problem = Problem([system], PenaltyLoss([obj], []))

    def make_batch(Bn, device, s=0):
        g = _gen(device, 8000 + s)
        return {"x": (0.1 + 0.8 * torch.rand(Bn, 1, 2, generator=g)).to(device),
                "r": torch.rand(Bn, nsteps + 1, 2, generator=g).to(device), "name": "train"}
    return Case("t_twotank_rk2_sigmoid", problem, system, nsteps, 130, make_batch, 0, 0)


def t_ssm_softplus(seed=0, nsteps=10) -> Case:
    torch.manual_seed(seed)
    nx, nu, nd, ny = 3, 2, 2, 2
    g0 = _gen("cpu", 99)
    A = 0.3 * torch.randn(nx, nx, generator=g0) + 0.5 * torch.eye(nx)
    Bm, E, Cm = (0.5 * torch.randn(*s, generator=g0) for s in ((nx, nu), (nx, nd), (ny, nx)))
    ssm = ode.SSM(_frozen_linear(A), _frozen_linear(Bm), nx, nu, fd=_frozen_linear(E), nd=nd)
    net = blocks.MLP(insize=ny + nd + nx, outsize=nu, hsizes=[10, 10], nonlin=nn.Softplus)
    system = System([Node(_frozen_linear(Cm), ["x"], ["y"], name="obs"),
                     Node(net, ["y", "d", "x"], ["u"], name="policy"),
                     Node(ssm, ["x", "u", "d"], ["x"], name="ssm")], nsteps=nsteps, name="sys")
    x, u, y = variable("x"), variable("u"), variable("y")
    o1 = (y == 1.) ^ 2
    o2 = 0.1 * ((u == 0.) ^ 2)
    c1_ = 5. * (x[:, 1:, 0:2] < 2.0)
    o1.name, o2.name, c1_.name = "o1", "o2", "c1"
    problem = Problem([system], PenaltyLoss([o1, o2], [c1_]))

    def make_batch(Bn, device, s=0):
        g = _gen(device, 9000 + s)
        return {"x": torch.randn(Bn, 1, nx, generator=g).to(device),
                "d": torch.randn(Bn, nsteps + 3, nd, generator=g).to(device), "name": "train"}
    return Case("t_ssm_softplus", problem, system, nsteps, 77, make_batch, 0, 0)


def t_node_policy_euler(seed=0, nsteps=6) -> Case:
    """Neural policy AND neural right-hand side with a control input (both trainable)."""
    torch.manual_seed(seed)
    nx, nu = 3, 2
    rhs = blocks.MLP(nx + nu, nx, bias=True, linear_map=nn.Linear, nonlin=nn.Tanh, hsizes=[12, 12])
    model = Node(integrators.RK2(rhs, h=0.2), ["x", "u"], ["x"], name="node")
    net = blocks.MLP(insize=nx, outsize=nu, hsizes=[8, 8], nonlin=nn.ReLU)
    system = System([Node(net, ["x"], ["u"], name="policy"), model], nsteps=nsteps, name="sys")
    x, u = variable("x"), variable("u")
    obj = (x == 0.5) ^ 2
    eff = 0.2 * ((u == 0.) ^ 2)
    obj.name, eff.name = "obj", "eff"
    problem = Problem([system], PenaltyLoss([obj, eff], []))

    def make_batch(Bn, device, s=0):
        g = _gen(device, 9500 + s)
        return {"x": torch.randn(Bn, 1, nx, generator=g).to(device), "name": "train"}
    return Case("t_node_policy_euler", problem, system, nsteps, 150, make_batch, 0, 0)


# ---- remaining explicit integrators and white-box right-hand sides --------------
def _freeze(module):
    for q in module.parameters():
        q.requires_grad_(False)
    return module


def t_int_euler_trap_lorenz(seed=0, nsteps=8) -> Case:
    torch.manual_seed(seed)
    model = Node(integrators.Euler_Trap(_freeze(ode.LorenzControl()), h=0.01), ["x", "u"], ["x"], name="model")
    net = blocks.MLP_bounds(insize=3, outsize=2, hsizes=[12], nonlin=nn.Tanh, min=-2., max=2., method="relu_clamp")
    system = System([Node(net, ["x"], ["u"], name="policy"), model], nsteps=nsteps, name="sys")
    obj = (variable("x") == 0.) ^ 2
    eff = 0.1 * ((variable("u") == 0.) ^ 2)
    obj.name, eff.name = "obj", "eff"
    problem = Problem([system], PenaltyLoss([obj, eff], []))
    mb = lambda Bn, device, s=0: {"x": (2.0 * torch.randn(Bn, 1, 3, generator=_gen(device, 9100 + s))).to(device), "name": "train"}
    return Case("t_int_euler_trap_lorenz", problem, system, nsteps, 100, mb, 0, 0)


def t_int_rk4_trap_cstr(seed=0, nsteps=6) -> Case:
    torch.manual_seed(seed)
    model = Node(integrators.RK4_Trap(_freeze(ode.CSTR_Param()), h=0.02), ["x", "u"], ["x"], name="model")
    net = blocks.MLP_bounds(insize=3, outsize=1, hsizes=[10, 10], nonlin=activations["gelu"], min=280., max=320.)
    system = System([Node(net, ["x", "r"], ["u"], name="policy"), model], nsteps=nsteps, name="sys")
    x, r = variable("x"), variable("r")
    obj = (x[:, :, [1]] == r) ^ 2
    cmax = 5. * (x[:, :, [0]] < 0.9)
    obj.name, cmax.name = "obj", "cmax"
    problem = Problem([system], PenaltyLoss([obj], [cmax]))

    def mb(Bn, device, s=0):
        g = _gen(device, 9200 + s)
        x0 = torch.cat([0.8 + 0.2 * torch.rand(Bn, 1, 1, generator=g), 300. + 20. * torch.rand(Bn, 1, 1, generator=g)], dim=-1)
        return {"x": x0.to(device), "r": (305. + 10. * torch.rand(Bn, nsteps + 1, 1, generator=g)).to(device), "name": "train"}
    return Case("t_int_rk4_trap_cstr", problem, system, nsteps, 100, mb, 0, 0)


def t_int_luther_vdp(seed=0, nsteps=7) -> Case:
    torch.manual_seed(seed)
    plant = ode.VanDerPolControl()
    plant.mu = nn.Parameter(torch.tensor(1.0), requires_grad=False)
    model = Node(integrators.Luther(plant, h=0.1), ["x", "u"], ["x"], name="model")
    net = blocks.MLP(insize=2, outsize=1, hsizes=[8, 8], nonlin=nn.ReLU)
    system = System([Node(net, ["x"], ["u"], name="policy"), model], nsteps=nsteps, name="sys")
    obj = (variable("x") == 0.) ^ 2
    obj.name = "obj"
    problem = Problem([system], PenaltyLoss([obj], []))
    mb = lambda Bn, device, s=0: {"x": torch.randn(Bn, 1, 2, generator=_gen(device, 9300 + s)).to(device), "name": "train"}
    return Case("t_int_luther_vdp", problem, system, nsteps, 100, mb, 0, 0)


def t_int_rkf_node(seed=0, nsteps=6) -> Case:
    """Open-loop neural ODE with a measured input read from the data (Part_4a_nonauto_NODE shape)."""
    torch.manual_seed(seed)
    fx = blocks.MLP(3, 2, bias=True, linear_map=nn.Linear, nonlin=nn.Tanh, hsizes=[16, 16])
    model = Node(integrators.Runge_Kutta_Fehlberg(fx, h=0.1), ["xn", "U"], ["xn"], name="NODE")
    system = System([model], name="system")                       # nsteps from data['X']
    ref = (variable("xn")[:, :-1, :] == variable("X")) ^ 2
    ref.name = "ref"
    problem = Problem([system], PenaltyLoss([ref], []))

    def mb(Bn, device, s=0):
        g = _gen(device, 9400 + s)
        X = torch.randn(Bn, nsteps, 2, generator=g)
        return {"X": X.to(device), "xn": X[:, 0:1, :].contiguous().to(device), "U": torch.randn(Bn, nsteps, 1, generator=g).to(device),
                "name": "train"}
    return Case("t_int_rkf_node", problem, system, nsteps, 100, mb, 0, 0)


def _t_second_order(integ, name, seed, nsteps) -> Case:
    """Second-order neural ODE ddx = MLP([x, u]) on X = [x, dx] with a measured input (reference integrators.py:340-404)."""
    torch.manual_seed(seed)
    fx = blocks.MLP(3, 2, bias=True, linear_map=nn.Linear, nonlin=nn.Tanh, hsizes=[16, 16])
    model = Node(integ(fx, h=0.1), ["xn", "U"], ["xn"], name="NODE")
    system = System([model], name="system")                       # nsteps from data['X']
    ref = (variable("xn")[:, :-1, :] == variable("X")) ^ 2
    ref.name = "ref"
    problem = Problem([system], PenaltyLoss([ref], []))

    def mb(Bn, device, s=0):
        g = _gen(device, 9700 + s)
        X = torch.randn(Bn, nsteps, 4, generator=g)
        return {"X": X.to(device), "xn": X[:, 0:1, :].contiguous().to(device), "U": torch.randn(Bn, nsteps, 1, generator=g).to(device),
                "name": "train"}
    return Case(name, problem, system, nsteps, 100, mb, 0, 0)


def t_int_leapfrog_node(seed=0, nsteps=7) -> Case:
    return _t_second_order(integrators.LeapFrog, "t_int_leapfrog_node", seed, nsteps)


def t_int_yoshida_node(seed=0, nsteps=7) -> Case:
    return _t_second_order(integrators.Yoshida4, "t_int_yoshida_node", seed, nsteps)


def t_open_duffing(seed=0, nsteps=9) -> Case:
    torch.manual_seed(seed)
    model = Node(integrators.RK4(_freeze(ode.DuffingParam()), h=0.05), ["x", "t"], ["x"], name="model")
    system = System([model], nsteps=nsteps, name="sys")
    obj = (variable("x") == 0.) ^ 2
    obj.name = "obj"
    problem = Problem([system], PenaltyLoss([obj], []))

    def mb(Bn, device, s=0):
        g = _gen(device, 9500 + s)
        t = (0.05 * torch.arange(nsteps, dtype=torch.float32)).reshape(1, nsteps, 1).repeat(Bn, 1, 1)
        return {"x": torch.randn(Bn, 1, 2, generator=g).to(device), "t": t.to(device), "name": "train"}
    return Case("t_open_duffing", problem, system, nsteps, 100, mb, 0, 0)


def _t_auto(name, plant, integ, h, nsteps, norm, target, seed, scale, shift):
    torch.manual_seed(seed)
    model = Node(integ(_freeze(plant), h=h), ["x"], ["x"], name="model")
    system = System([model], nsteps=nsteps, name="sys")
    obj = (variable("x") == target) ^ norm
    obj.name = "obj"
    problem = Problem([system], PenaltyLoss([obj], []))
    nx = plant.out_features
    mb = lambda Bn, device, s=0: {"x": (shift + scale * torch.rand(Bn, 1, nx, generator=_gen(device, 9600 + s))).to(device), "name": "train"}
    return Case(name, problem, system, nsteps, 100, mb, 0, 0)


def t_auto_brusselator(seed=0, nsteps=8) -> Case:
    plant = ode.BrusselatorParam()
    plant.alpha = nn.Parameter(torch.tensor([1.0]), requires_grad=False)
    plant.beta = nn.Parameter(torch.tensor([3.0]), requires_grad=False)
    return _t_auto("t_auto_brusselator", plant, integrators.RK2, 0.05, nsteps, 1, 1., seed, 2.0, 0.5)


def t_auto_lotka(seed=0, nsteps=8) -> Case:
    return _t_auto("t_auto_lotka", ode.LotkaVolterraParam(), integrators.Euler, 0.1, nsteps, 2, 1., seed, 3.0, 0.5)


def t_auto_lorenz(seed=0, nsteps=8) -> Case:
    plant = ode.LorenzParam()
    plant.rho = nn.Parameter(torch.tensor([28.0]), requires_grad=False)
    plant.sigma = nn.Parameter(torch.tensor([10.0]), requires_grad=False)
    plant.beta = nn.Parameter(torch.tensor([2.66667]), requires_grad=False)
    return _t_auto("t_auto_lorenz", plant, integrators.RK4, 0.01, nsteps, 2, 0., seed, 10.0, -5.0)


# ---- system identification: gradients w.r.t. the white-box constants (reference ode.py declares them
# trainable nn.Parameters, examples/ODEs/Part_1_param_estim_ODE / Part_2_param_estim_ODE fit them to measured trajectories)
def _t_sysid(name, plant, integ, h, nsteps, seed, x0_fn, u_fn=None, policy=None, norm=2) -> Case:
    torch.manual_seed(seed)
    nx = plant.out_features
    if policy is not None:                       # closed loop: policy and plant constants are trained together
        model = Node(integ(plant, h=h), ["x", "u"], ["x"], name="model")
        system = System([Node(policy, ["x"], ["u"], name="policy"), model], nsteps=nsteps, name="sys")
        obj = (variable("x") == 0.) ^ 2
        obj.name = "obj"
        problem = Problem([system], PenaltyLoss([obj], []))
        mb = lambda Bn, device, s=0: {"x": x0_fn(Bn, _gen(device, 9800 + s)).to(device), "name": "train"}
        return Case(name, problem, system, nsteps, 100, mb, 0, 0)
    keys = ["xn"] + (["U"] if u_fn is not None else [])
    model = Node(integ(plant, h=h), keys, ["xn"], name="model")
    system = System([model], name="system")                       # nsteps from data['X']
    ref = (variable("xn")[:, :-1, :] == variable("X")) ^ norm
    ref.name = "ref"
    problem = Problem([system], PenaltyLoss([ref], []))

    def mb(Bn, device, s=0):
        g = _gen(device, 9800 + s)
        x0 = x0_fn(Bn, g)
        X = x0 + 0.1 * torch.randn(Bn, nsteps, nx, generator=g)    # "measured" trajectory around the initial condition
        out = {"X": X.to(device), "xn": X[:, 0:1, :].contiguous().to(device), "name": "train"}
        if u_fn is not None:
            out["U"] = u_fn(Bn, g).to(device)
        return out
    return Case(name, problem, system, nsteps, 100, mb, 0, 0)


def t_sysid_twotank(seed=0, nsteps=8) -> Case:
    plant = ode.TwoTankParam()
    plant.c1, plant.c2 = nn.Parameter(torch.tensor([0.08])), nn.Parameter(torch.tensor([0.04]))
    return _t_sysid("t_sysid_twotank", plant, integrators.RK4, 1.0, nsteps, seed, lambda B, g: 0.2 + 0.5 * torch.rand(B, 1, 2, generator=g),
                    u_fn=lambda B, g: torch.rand(B, nsteps, 2, generator=g))


def t_sysid_lorenz(seed=0, nsteps=8) -> Case:
    plant = ode.LorenzParam()
    plant.rho, plant.sigma, plant.beta = (nn.Parameter(torch.tensor([v])) for v in (20.0, 8.0, 2.0))
    return _t_sysid("t_sysid_lorenz", plant, integrators.RK4, 0.01, nsteps, seed, lambda B, g: -5.0 + 10.0 * torch.rand(B, 1, 3, generator=g))


def t_sysid_cstr(seed=0, nsteps=6) -> Case:
    """all ten CSTR constants trainable (they are in the reference): exercises the derived coefficient groups q/V, mdelH/(rho Cp), UA/V/rho/Cp"""
    base = lambda B, g: torch.tensor([0.5, 330.0]) + torch.tensor([0.4, 20.0]) * torch.rand(B, 1, 2, generator=g)
    return _t_sysid("t_sysid_cstr", ode.CSTR_Param(), integrators.Euler, 0.01, nsteps, seed, base,
                    u_fn=lambda B, g: 290.0 + 20.0 * torch.rand(B, nsteps, 1, generator=g))


def t_sysid_duffing(seed=0, nsteps=9) -> Case:
    plant = ode.DuffingParam()
    for q in plant.parameters():
        q.requires_grad_(True)
    tt = lambda B, g: (0.05 * torch.arange(nsteps, dtype=torch.float32)).reshape(1, nsteps, 1).repeat(B, 1, 1)
    return _t_sysid("t_sysid_duffing", plant, integrators.RK2, 0.05, nsteps, seed, lambda B, g: torch.randn(B, 1, 2, generator=g), u_fn=tt)


def t_sysid_brusselator(seed=0, nsteps=8) -> Case:
    plant = ode.BrusselatorParam()
    plant.alpha, plant.beta = nn.Parameter(torch.tensor([1.0])), nn.Parameter(torch.tensor([3.0]))
    return _t_sysid("t_sysid_brusselator", plant, integrators.RK2, 0.05, nsteps, seed, lambda B, g: 0.5 + 2.0 * torch.rand(B, 1, 2, generator=g), norm=1)


def t_sysid_lotka(seed=0, nsteps=8) -> Case:
    return _t_sysid("t_sysid_lotka", ode.LotkaVolterraParam(), integrators.Euler, 0.1, nsteps, seed, lambda B, g: 0.5 + 3.0 * torch.rand(B, 1, 2, generator=g))


def t_ude_brusselator(seed=0, nsteps=6) -> Case:
    """Universal differential equation (examples/ODEs/Part_3a_UDE.py:94-104): the Brusselator's x2*x1^2 term is an MLP."""
    torch.manual_seed(seed)
    net = blocks.MLP(2, 1, bias=True, linear_map=nn.Linear, nonlin=nn.GELU, hsizes=[20, 20])
    return _t_sysid("t_ude_brusselator", ode.BrusselatorHybrid(net), integrators.RK4, 0.05, nsteps, seed, lambda B, g: 0.5 + 2.0 * torch.rand(B, 1, 2, generator=g))


def t_ude_lotka(seed=0, nsteps=6) -> Case:
    torch.manual_seed(seed)
    net = blocks.MLP(2, 1, bias=True, linear_map=nn.Linear, nonlin=nn.Tanh, hsizes=[16])
    return _t_sysid("t_ude_lotka", ode.LotkaVolterraHybrid(net), integrators.Euler, 0.1, nsteps, seed, lambda B, g: 0.5 + 3.0 * torch.rand(B, 1, 2, generator=g))


def t_sysid_vdp_policy(seed=0, nsteps=10) -> Case:
    torch.manual_seed(seed)
    net = blocks.MLP(insize=2, outsize=1, hsizes=[16, 16], bias=True, linear_map=nn.Linear, nonlin=nn.Tanh)
    return _t_sysid("t_sysid_vdp_policy", ode.VanDerPolControl(), integrators.RK4, 0.1, nsteps, seed, lambda B, g: torch.randn(B, 1, 2, generator=g), policy=net)


def _t_preview(pad_mode, seed=0, nsteps=9) -> Case:
    """TwoTank tracking with a previewing policy (examples/control/Part_3_ref_tracking_ODE.py:93-105)."""
    torch.manual_seed(seed)
    plant = ode.TwoTankParam()
    plant.c1 = nn.Parameter(torch.tensor(0.08), requires_grad=False)
    plant.c2 = nn.Parameter(torch.tensor(0.04), requires_grad=False)
    model = Node(integrators.RK4(plant, h=1.0), ["x", "u"], ["x"], name="model")
    net = blocks.MLP_bounds(insize=2 + 2 * 6 + 1 * 3, outsize=2, hsizes=[16], nonlin=activations["gelu"], min=0., max=1.)
    policy = Node(net, ["x", "r", "p"], ["u"], name="policy")
    system = SystemPreview([policy, model], name="sys", nsteps=nsteps, preview_keys_map={"r": ["policy"], "p": ["policy"]},
                           preview_length={"r": 5, "p": 2}, pad_mode=pad_mode, pad_constant=0.25)
    x, rr = variable("x"), variable("r")
    track = 5. * ((x[:, 1:, :] == rr[:, :nsteps, :]) ^ 2)
    lo = 10. * (x > 0.)
    track.name, lo.name = "track", "lo"
    problem = Problem([system], PenaltyLoss([track], [lo]))

    def make_batch(Bn, device, s=0):
        g = _gen(device, 9700 + s)
        return {"x": torch.rand(Bn, 1, 2, generator=g).to(device), "r": torch.rand(Bn, nsteps + 2, 2, generator=g).to(device),
                "p": torch.rand(Bn, nsteps, 1, generator=g).to(device), "name": "train"}
    return Case(f"t_preview_{pad_mode}", problem, system, nsteps, 140, make_batch, 0, 0)


def t_preview_replicate(seed=0, **kw): return _t_preview("replicate", seed, **kw)
def t_preview_circular(seed=0, **kw): return _t_preview("circular", seed, **kw)
def t_preview_constant(seed=0, **kw): return _t_preview("constant", seed, **kw)
def t_preview_reflect(seed=0, **kw): return _t_preview("reflect", seed, **kw)


def t_preview_index(seed=0, nsteps=7, pad_mode="replicate") -> Case:
    """Indexing probe: a bias-free one-layer policy whose weight rows are one-hot, feeding x+ = u.  The
    state trajectory then consists of *copies* of preview-window entries, so the window / padding index
    arithmetic can be checked bit-exactly (tests/test_parity_gpu.py)."""
    torch.manual_seed(seed)
    nx = nu = 2
    L_s, d_s = 3, 2                                    # signal `s`: window 4, width 2 -> 8 policy inputs
    net = blocks.MLP(insize=(L_s + 1) * d_s, outsize=nu, hsizes=[], bias=False, linear_map=nn.Linear, nonlin=nn.ReLU)
    ssm = ode.SSM(_frozen_linear(torch.zeros(nx, nx)), _frozen_linear(torch.eye(nx)), nx, nu)
    system = SystemPreview([Node(net, ["s"], ["u"], name="policy"), Node(ssm, ["x", "u"], ["x"], name="plant")],
                           name="sys", nsteps=nsteps, preview_keys_map={"s": ["policy"]}, preview_length={"s": L_s},
                           pad_mode=pad_mode, pad_constant=-7.5)
    obj = (variable("x") == 0.) ^ 2
    obj.name = "obj"
    problem = Problem([system], PenaltyLoss([obj], []))

    def make_batch(Bn, device, s=0):
        g = _gen(device, 9900 + s)
        return {"x": torch.zeros(Bn, 1, nx, device=device), "s": torch.randn(Bn, nsteps, d_s, generator=g).to(device),
                "name": "train"}
    return Case("t_preview_index", problem, system, nsteps, 64, make_batch, 0, 0)


CASES: Dict[str, Callable[..., Case]] = {
    "c1": c1, "c2": c2, "c3": c3, "c4": c4, "c5": c5,
    "t_vdp_rk4_tanh_clamp": t_vdp_rk4_tanh_clamp, "t_vdp_euler_elu": t_vdp_euler_elu,
    "t_twotank_rk2_sigmoid": t_twotank_rk2_sigmoid, "t_ssm_softplus": t_ssm_softplus,
    "t_node_policy_euler": t_node_policy_euler,
    "t_int_euler_trap_lorenz": t_int_euler_trap_lorenz, "t_int_rk4_trap_cstr": t_int_rk4_trap_cstr,
    "t_int_luther_vdp": t_int_luther_vdp, "t_int_rkf_node": t_int_rkf_node, "t_open_duffing": t_open_duffing,
    "t_auto_brusselator": t_auto_brusselator, "t_auto_lotka": t_auto_lotka, "t_auto_lorenz": t_auto_lorenz,
    "t_preview_replicate": t_preview_replicate, "t_preview_circular": t_preview_circular,
    "t_preview_constant": t_preview_constant, "t_preview_reflect": t_preview_reflect,
    "t_slim_pf": t_slim_pf, "t_slim_svd": t_slim_svd,
    "t_int_leapfrog_node": t_int_leapfrog_node, "t_int_yoshida_node": t_int_yoshida_node,
    "t_sysid_twotank": t_sysid_twotank, "t_sysid_lorenz": t_sysid_lorenz, "t_sysid_cstr": t_sysid_cstr, "t_sysid_duffing": t_sysid_duffing,
    "t_sysid_brusselator": t_sysid_brusselator, "t_sysid_lotka": t_sysid_lotka, "t_sysid_vdp_policy": t_sysid_vdp_policy,
    "t_ude_brusselator": t_ude_brusselator, "t_ude_lotka": t_ude_lotka,
    "t_barrier_log10": t_barrier_log10, "t_barrier_log": t_barrier_log, "t_barrier_inverse": t_barrier_inverse,
    "t_barrier_softexp": t_barrier_softexp, "t_barrier_softlog": t_barrier_softlog, "t_barrier_expshift": t_barrier_expshift,
}
# shape classes compiled ahead of time that have no oracle / golden counterpart (special-purpose probes)
EXTRA_CASES: Dict[str, Callable[..., Case]] = {"t_preview_index": t_preview_index, "c4p": c4p}
AOT_CASES = list(CASES) + list(EXTRA_CASES)


def build_case(name: str, batch=None, device="cpu", seed=0, **kw) -> Case:
    return {**CASES, **EXTRA_CASES}[name](seed=seed, **kw)


def make_case(name: str, device, seed=0, **kw) -> Case:
    """``build_case`` moved to ``device`` with its lowered plan parameters attached (``case.plan_params``: the
    parameters in the order of the flat arena) -- what bench.py and the parity tests run."""
    case = build_case(name, seed=seed, **kw)
    case.problem.to(device)
    plan = case.plan()
    # a structured slim map stands in the plan as its materialised matrix: the tests address its own parameters
    # (registration order, the bias -- a plan entry of its own -- left out)
    case.plan_params = []
    for q in plan.params:
        if hasattr(q, "materialise"):
            case.plan_params += q.own_parameters()
        else:
            case.plan_params.append(q)
    return case
