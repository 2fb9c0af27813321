"""
TEST INFRASTRUCTURE -- oracle-side definitions of the benchmark / test problems.

Each builder returns the ``spec`` consumed by ``dpc_oracle.run_case`` for one named case, given the
flat list of trainable tensors (W0, b0, W1, b1, ... of the policy, then of the MLP right-hand side).
The problems are the ones of ``BASELINE.json`` as the reference examples define them (file:line in
each builder); constants are restated here independently of the product package.
"""
from __future__ import annotations

from typing import Dict, List

import torch

from . import dpc_oracle as O


def _term(name, objective, cmp, norm, weight, left, right):
    return dict(name=name, objective=objective, cmp=cmp, norm=norm, weight=weight, left=left, right=right)


def _t(v, like):
    return torch.tensor(v, dtype=like.dtype)


# c1 -- examples/control/Part_1_stabilize_linear_system.py:14-67 (horizon 50 per BASELINE.json)
def c1(params: List[torch.Tensor], nsteps=50) -> dict:
    A = _t([[1.2, 1.0], [0.0, 1.0]], params[0])
    B = _t([[1.0], [0.5]], params[0])
    nodes = [dict(fn=lambda x: O.mlp_forward(params, x, "relu"), inputs=["X"], outputs=["U"]),
             dict(fn=lambda x, u: O.ssm_step(x, u, None, A, B), inputs=["X", "U"], outputs=["X"])]
    terms = [_term("action_loss", True, "eq", 2, 0.0001, lambda d: d["U"], lambda d: 0.),
             _term("regulation_loss", True, "eq", 2, 10., lambda d: d["X"], lambda d: 0.)]
    return dict(nodes=nodes, terms=terms, nsteps=nsteps, params=params, traj_keys=["X", "U"])


# c2 -- examples/control/Part_3_ref_tracking_ODE.py:32-163
def c2(params, nsteps=30) -> dict:
    c1_, c2_ = _t(0.08, params[0]), _t(0.04, params[0])
    h = _t(1.0, params[0])
    rhs = lambda x, u: O.rhs_twotank(x, u, c1_, c2_)
    nodes = [dict(fn=lambda x, r: O.bounds_scaling(O.mlp_forward(params, torch.cat([x, r], dim=-1), "gelu"), 0, 1.),
                  inputs=["x", "r"], outputs=["u"]),
             dict(fn=lambda x, u: O.rk4_step(rhs, x, h, u), inputs=["x", "u"], outputs=["x"])]
    terms = [_term("ref_tracking", True, "eq", 2, 5., lambda d: d["x"], lambda d: d["r"]),
             _term("x_min", False, "gt", 1, 10., lambda d: d["x"], lambda d: 0),
             _term("x_max", False, "lt", 1, 10., lambda d: d["x"], lambda d: 1.),
             _term("x_N_min", False, "gt", 1, 10., lambda d: d["x"][:, [-1], :], lambda d: d["r"] - 0.01),
             _term("x_N_max", False, "lt", 1, 10., lambda d: d["x"][:, [-1], :], lambda d: d["r"] + 0.01)]
    return dict(nodes=nodes, terms=terms, nsteps=nsteps, params=params, traj_keys=["x", "u"])


# c3 -- examples/ODEs/Part_1_NODE.py:58-104 with integrators.RK4
def c3(params, nsteps=100) -> dict:
    f = lambda x: O.mlp_forward(params, x, "relu")
    nodes = [dict(fn=lambda x: O.rk4_step(f, x, 0.1), inputs=["xn"], outputs=["xn"])]
    xhat = lambda d: d["xn"][:, :-1, :]
    terms = [_term("ref_loss", True, "eq", 2, 1.0, xhat, lambda d: d["X"]),
             _term("FD_loss", True, "eq", 2, 2., lambda d: d["X"][:, 1:, :] - d["X"][:, :-1, :],
                   lambda d: xhat(d)[:, 1:, :] - xhat(d)[:, :-1, :])]
    return dict(nodes=nodes, terms=terms, nsteps=nsteps, params=params, traj_keys=["xn"])


# c4 -- examples/domain_examples/grid_response.ipynb cells 7-30 with synthetic Schur-stable matrices
C4_A = [[0.9500, 0.0200, 0.0000, 0.0100], [0.0300, 0.9000, 0.0200, 0.0000],
        [0.0000, 0.0400, 0.9200, 0.0100], [0.0200, 0.0100, 0.0300, 0.9300]]
C4_B = [[0.00], [0.00], [0.05], [0.10]]
C4_E = [[0.010, 0.000, 0.002], [0.020, 0.001, 0.000], [0.005, 0.002, 0.001], [0.002, 0.001, 0.003]]
C4_C = [[0.0, 0.0, 0.0, 1.0]]


def c4(params, nsteps=96) -> dict:
    A, B, E, C = (_t(m, params[0]) for m in (C4_A, C4_B, C4_E, C4_C))
    umin, umax = _t([0.0], params[0]), _t([5000.0], params[0])
    keys = ["y", "d_obs_lookahead", "ymin_lookahead", "ymax_lookahead", "price_lookahead"]
    nodes = [dict(fn=lambda x: torch.nn.functional.linear(x, C), inputs=["x"], outputs=["y"]),
             dict(fn=lambda *a: O.bounds_scaling(O.mlp_forward(params, torch.cat(a, dim=-1), "leakyrelu"), umin, umax),
                  inputs=keys, outputs=["u"]),
             dict(fn=lambda x, u, d: O.ssm_step(x, u, d, A, B, E), inputs=["x", "u", "d"], outputs=["x"])]
    terms = [_term("price_loss", True, "eq", 2, 1, lambda d: d["u"] * d["price_lookahead"][:, :-1, 0:1],
                   lambda d: torch.zeros_like(d["u"])),
             _term("y_min", False, "gt", 1, 5e06, lambda d: d["y"], lambda d: d["ymin_lookahead"][:, :-1, 0:1]),
             _term("y_max", False, "lt", 1, 5e06, lambda d: d["y"], lambda d: d["ymax_lookahead"][:, :-1, 0:1])]
    return dict(nodes=nodes, terms=terms, nsteps=nsteps, params=params, traj_keys=["x", "u", "y"])


# c5 -- examples/control/Part_2_stabilize_ODE.py:33-135 scaled to 6 coupled oscillators
def c5(params, nsteps=200) -> dict:
    mu, kappa, h = _t([1.0], params[0]), _t([0.1], params[0]), _t(0.05, params[0])
    rhs = lambda x, u: O.rhs_cvdp(x, u, mu, kappa)
    nodes = [dict(fn=lambda x, r: O.bounds_scaling(O.mlp_forward(params, torch.cat([x, r], dim=-1), "gelu"), -5., 5.),
                  inputs=["x", "r"], outputs=["u"]),
             dict(fn=lambda x, u: O.rk4_step(rhs, x, h, u), inputs=["x", "u"], outputs=["x"])]
    terms = [_term("state_loss", True, "eq", 2, 1., lambda d: d["x"], lambda d: d["r"]),
             _term("action_loss", True, "eq", 2, 0.01, lambda d: d["u"], lambda d: 0.),
             _term("x_min", False, "gt", 1, 10., lambda d: d["x"], lambda d: -4.),
             _term("x_max", False, "lt", 1, 10., lambda d: d["x"], lambda d: 4.),
             _term("x_N_min", False, "gt", 1, 10., lambda d: d["x"][:, [-1], :], lambda d: d["r"] - 0.01),
             _term("x_N_max", False, "lt", 1, 10., lambda d: d["x"][:, [-1], :], lambda d: d["r"] + 0.01)]
    return dict(nodes=nodes, terms=terms, nsteps=nsteps, params=params, traj_keys=["x", "u"])


# ---- small shape classes exercising the remaining activations / bounds / integrators ------------
def t_vdp_rk4_tanh_clamp(params, nsteps=12) -> dict:
    mu = _t(1.0, params[0])
    rhs = lambda x, u: O.rhs_vdp(x, u, mu)
    nodes = [dict(fn=lambda x, r: O.bounds_clamp(O.mlp_forward(params, torch.cat([x, r], dim=-1), "tanh"), -1., 1.),
                  inputs=["x", "r"], outputs=["u"]),
             dict(fn=lambda x, u: O.rk4_step(rhs, x, 0.1, u), inputs=["x", "u"], outputs=["x"])]
    terms = [_term("obj", True, "eq", 2, 3., lambda d: d["x"], lambda d: d["r"]),
             _term("eff", True, "eq", 1, 0.1, lambda d: d["u"], lambda d: 0.),
             _term("du", True, "eq", 2, 0.5, lambda d: d["u"][:, 1:, :] - d["u"][:, :-1, :], lambda d: 0.),
             _term("box", False, "lt", 2, 2., lambda d: d["x"], lambda d: 1.5)]
    return dict(nodes=nodes, terms=terms, nsteps=nsteps, params=params, traj_keys=["x", "u"])


# BarrierLoss (reference loss.py:184-257) on the tanh-clamp problem: box constraints on the state (norm 2
# upper, norm 1 lower) and on the action, every built-in barrier function
def _t_barrier(params, nsteps, barrier):
    spec = t_vdp_rk4_tanh_clamp(params, nsteps)
    spec["terms"] = [_term("obj", True, "eq", 2, 3., lambda d: d["x"], lambda d: d["r"]),
                     _term("hi", False, "lt", 2, 2., lambda d: d["x"], lambda d: 1.5),
                     _term("lo", False, "gt", 1, 1., lambda d: d["x"], lambda d: -1.5),
                     _term("umax", False, "lt", 1, 0.5, lambda d: d["u"], lambda d: 0.8)]
    # (expshift: the reference's own backward raises on its in-place masked assignment -- see dpc_oracle.barrier_loss)
    spec["barrier"] = dict(barrier=barrier, upper_bound=2.0, shift=0.7, alpha=0.4, out_of_place=barrier == "expshift")
    return spec


def t_barrier_log10(params, nsteps=12): return _t_barrier(params, nsteps, "log10")
def t_barrier_log(params, nsteps=12): return _t_barrier(params, nsteps, "log")
def t_barrier_inverse(params, nsteps=12): return _t_barrier(params, nsteps, "inverse")
def t_barrier_softexp(params, nsteps=12): return _t_barrier(params, nsteps, "softexp")
def t_barrier_softlog(params, nsteps=12): return _t_barrier(params, nsteps, "softlog")
def t_barrier_expshift(params, nsteps=12): return _t_barrier(params, nsteps, "expshift")


# structured slim maps: the tanh-clamp problem whose policy layers are slim maps; `params` holds, per layer,
# the map's own parameters (registration order) and then its bias [1, out]
def _slim_case(params, nsteps, per_layer, eff_w):
    n_layers = len(params) // (per_layer + 1)

    def policy(z):                                          # LinearBase.forward: x @ effective_W() + bias (slim/linear.py:94-101)
        for l in range(n_layers):
            chunk = params[l * (per_layer + 1):(l + 1) * (per_layer + 1)]
            z = torch.matmul(z, eff_w(*chunk[:per_layer])) + chunk[per_layer]
            if l < n_layers - 1:
                z = torch.tanh(z)
        return z
    spec = t_vdp_rk4_tanh_clamp(params, nsteps=nsteps)
    spec["nodes"][0] = dict(fn=lambda x, r: O.bounds_clamp(policy(torch.cat([x, r], dim=-1)), -1., 1.), inputs=["x", "r"], outputs=["u"])
    return spec


def t_slim_pf(params, nsteps=12) -> dict:
    return _slim_case(params, nsteps, 2, lambda w, s: O.slim_perron_frobenius(w, s, 0.8, 1.0))


def t_slim_svd(params, nsteps=12) -> dict:
    return _slim_case(params, nsteps, 3, lambda u, v, s: O.slim_soft_svd(u, v, s, 0.1, 1.0))


def t_vdp_euler_elu(params, nsteps=9) -> dict:
    mu = _t(0.7, params[0])
    rhs = lambda x, u: O.rhs_vdp(x, u, mu)
    nodes = [dict(fn=lambda x: O.mlp_forward(params, x, "elu", bias=False), inputs=["x"], outputs=["u"]),
             dict(fn=lambda x, u: O.euler_step(rhs, x, 0.05, u), inputs=["x", "u"], outputs=["x"])]
    terms = [_term("obj", True, "eq", 2, 1.0, lambda d: d["x"], lambda d: 0.),
             _term("pen", False, "gt", 1, 4., lambda d: d["u"], lambda d: -0.2)]
    return dict(nodes=nodes, terms=terms, nsteps=nsteps, params=params, traj_keys=["x", "u"])


def t_twotank_rk2_sigmoid(params, nsteps=7) -> dict:
    c1_, c2_ = _t(0.08, params[0]), _t(0.04, params[0])
    rhs = lambda x, u: O.rhs_twotank(x, u, c1_, c2_)
    nodes = [dict(fn=lambda x, r: O.bounds_scaling(O.mlp_forward(params, torch.cat([x, r], dim=-1), "sigmoid"), 0., 1.),
                  inputs=["x", "r"], outputs=["u"]),
             dict(fn=lambda x, u: O.rk2_step(rhs, x, 0.5, u), inputs=["x", "u"], outputs=["x"])]
    terms = [_term("obj", True, "eq", 2, 1.0, lambda d: d["x"], lambda d: d["r"])]
    return dict(nodes=nodes, terms=terms, nsteps=nsteps, params=params, traj_keys=["x", "u"])


def _ssm_softplus_mats(like):
    g0 = torch.Generator(device="cpu")
    g0.manual_seed(99)
    nx, nu, nd, ny = 3, 2, 2, 2
    A = 0.3 * torch.randn(nx, nx, generator=g0) + 0.5 * torch.eye(nx)
    B, E, C = (0.5 * torch.randn(*s, generator=g0) for s in ((nx, nu), (nx, nd), (ny, nx)))
    return [m.to(like.dtype) for m in (A, B, E, C)]


def t_ssm_softplus(params, nsteps=10) -> dict:
    A, B, E, C = _ssm_softplus_mats(params[0])
    nodes = [dict(fn=lambda x: torch.nn.functional.linear(x, C), inputs=["x"], outputs=["y"]),
             dict(fn=lambda y, d, x: O.mlp_forward(params, torch.cat([y, d, x], dim=-1), "softplus"),
                  inputs=["y", "d", "x"], outputs=["u"]),
             dict(fn=lambda x, u, d: O.ssm_step(x, u, d, A, B, E), inputs=["x", "u", "d"], outputs=["x"])]
    terms = [_term("o1", True, "eq", 2, 1.0, lambda d: d["y"], lambda d: 1.),
             _term("o2", True, "eq", 2, 0.1, lambda d: d["u"], lambda d: 0.),
             _term("c1", False, "lt", 1, 5., lambda d: d["x"][:, 1:, 0:2], lambda d: 2.0)]
    return dict(nodes=nodes, terms=terms, nsteps=nsteps, params=params, traj_keys=["x", "u", "y"])


def t_node_policy_euler(params, nsteps=6) -> dict:
    pol, rhs_p = params[:6], params[6:]
    f = lambda x, u: O.mlp_forward(rhs_p, torch.cat([x, u], dim=-1), "tanh")
    nodes = [dict(fn=lambda x: O.mlp_forward(pol, x, "relu"), inputs=["x"], outputs=["u"]),
             dict(fn=lambda x, u: O.rk2_step(f, x, 0.2, u), inputs=["x", "u"], outputs=["x"])]
    terms = [_term("obj", True, "eq", 2, 1.0, lambda d: d["x"], lambda d: 0.5),
             _term("eff", True, "eq", 2, 0.2, lambda d: d["u"], lambda d: 0.)]
    return dict(nodes=nodes, terms=terms, nsteps=nsteps, params=params, traj_keys=["x", "u"])


# ---- remaining explicit integrators and white-box right-hand sides --------------
def _cstr_consts(like):
    vals = dict(q=100.0, V=100.0, rho=1000.0, Cp=0.239, mdelH=5e4, EoverR=8750.0, k0=7.2e10, UA=5e4, Tf=350.0, Caf=1.0)
    return {k: _t([v], like) for k, v in vals.items()}


def t_int_euler_trap_lorenz(params, nsteps=8) -> dict:
    rho, sigma, beta = (_t([v], params[0]) for v in (28.0, 10.0, 2.66667))
    rhs = lambda x, u: O.rhs_lorenz_control(x, u, rho, sigma, beta)
    nodes = [dict(fn=lambda x: O.bounds_clamp(O.mlp_forward(params, x, "tanh"), -2., 2.), inputs=["x"], outputs=["u"]),
             dict(fn=lambda x, u: O.euler_trap_step(rhs, x, 0.01, u), inputs=["x", "u"], outputs=["x"])]
    terms = [_term("obj", True, "eq", 2, 1.0, lambda d: d["x"], lambda d: 0.),
             _term("eff", True, "eq", 2, 0.1, lambda d: d["u"], lambda d: 0.)]
    return dict(nodes=nodes, terms=terms, nsteps=nsteps, params=params, traj_keys=["x", "u"])


def t_int_rk4_trap_cstr(params, nsteps=6) -> dict:
    cc = _cstr_consts(params[0])
    rhs = lambda x, u: O.rhs_cstr(x, u, cc)
    nodes = [dict(fn=lambda x, r: O.bounds_scaling(O.mlp_forward(params, torch.cat([x, r], dim=-1), "gelu"), 280., 320.),
                  inputs=["x", "r"], outputs=["u"]),
             dict(fn=lambda x, u: O.rk4_trap_step(rhs, x, 0.02, u), inputs=["x", "u"], outputs=["x"])]
    terms = [_term("obj", True, "eq", 2, 1.0, lambda d: d["x"][:, :, [1]], lambda d: d["r"]),
             _term("cmax", False, "lt", 1, 5., lambda d: d["x"][:, :, [0]], lambda d: 0.9)]
    return dict(nodes=nodes, terms=terms, nsteps=nsteps, params=params, traj_keys=["x", "u"])


def t_int_luther_vdp(params, nsteps=7) -> dict:
    mu = _t(1.0, params[0])
    rhs = lambda x, u: O.rhs_vdp(x, u, mu)
    nodes = [dict(fn=lambda x: O.mlp_forward(params, x, "relu"), inputs=["x"], outputs=["u"]),
             dict(fn=lambda x, u: O.luther_step(rhs, x, 0.1, u), inputs=["x", "u"], outputs=["x"])]
    terms = [_term("obj", True, "eq", 2, 1.0, lambda d: d["x"], lambda d: 0.)]
    return dict(nodes=nodes, terms=terms, nsteps=nsteps, params=params, traj_keys=["x", "u"])


def t_int_rkf_node(params, nsteps=6) -> dict:
    """open-loop neural ODE with a measured input (examples/ODEs/Part_4a_nonauto_NODE.py:79-111 shape)"""
    f = lambda x, u: O.mlp_forward(params, torch.cat([x, u], dim=-1), "tanh")
    nodes = [dict(fn=lambda x, u: O.rkf_step(f, x, 0.1, u), inputs=["xn", "U"], outputs=["xn"])]
    terms = [_term("ref", True, "eq", 2, 1.0, lambda d: d["xn"][:, :-1, :], lambda d: d["X"])]
    return dict(nodes=nodes, terms=terms, nsteps=nsteps, params=params, traj_keys=["xn"])


def _t_second_order(params, nsteps, step):
    """second-order neural ODE ddx = MLP([x, u]) on X = [x, dx] with a measured input (integrators.py:340-404)"""
    f = lambda x, u: O.mlp_forward(params, torch.cat([x, u], dim=-1), "tanh")
    nodes = [dict(fn=lambda X, u: step(f, X, 0.1, u), inputs=["xn", "U"], outputs=["xn"])]
    terms = [_term("ref", True, "eq", 2, 1.0, lambda d: d["xn"][:, :-1, :], lambda d: d["X"])]
    return dict(nodes=nodes, terms=terms, nsteps=nsteps, params=params, traj_keys=["xn"])


def t_int_leapfrog_node(params, nsteps=7) -> dict:
    return _t_second_order(params, nsteps, O.leapfrog_step)


def t_int_yoshida_node(params, nsteps=7) -> dict:
    return _t_second_order(params, nsteps, O.yoshida4_step)


def t_open_duffing(params, nsteps=9) -> dict:
    a, b, dl, g, w = (_t([v], torch.zeros(1)) for v in (1.0, 5.0, 0.02, 8.0, 0.5))
    rhs = lambda x, t: O.rhs_duffing(x, t, a, b, dl, g, w)
    nodes = [dict(fn=lambda x, t: O.rk4_step(rhs, x, 0.05, t), inputs=["x", "t"], outputs=["x"])]
    terms = [_term("obj", True, "eq", 2, 1.0, lambda d: d["x"], lambda d: 0.)]
    return dict(nodes=nodes, terms=terms, nsteps=nsteps, params=params, traj_keys=["x"])


def t_auto_brusselator(params, nsteps=8) -> dict:
    al, be = _t([1.0], torch.zeros(1)), _t([3.0], torch.zeros(1))
    rhs = lambda x: O.rhs_brusselator(x, al, be)
    nodes = [dict(fn=lambda x: O.rk2_step(rhs, x, 0.05), inputs=["x"], outputs=["x"])]
    terms = [_term("obj", True, "eq", 1, 1.0, lambda d: d["x"], lambda d: 1.)]
    return dict(nodes=nodes, terms=terms, nsteps=nsteps, params=params, traj_keys=["x"])


def t_auto_lotka(params, nsteps=8) -> dict:
    al = _t([1.0], torch.zeros(1))
    rhs = lambda x: O.rhs_lotka_volterra(x, al)
    nodes = [dict(fn=lambda x: O.euler_step(rhs, x, 0.1), inputs=["x"], outputs=["x"])]
    terms = [_term("obj", True, "eq", 2, 1.0, lambda d: d["x"], lambda d: 1.)]
    return dict(nodes=nodes, terms=terms, nsteps=nsteps, params=params, traj_keys=["x"])


def t_auto_lorenz(params, nsteps=8) -> dict:
    rho, sigma, beta = (_t([v], torch.zeros(1)) for v in (28.0, 10.0, 2.66667))
    rhs = lambda x: O.rhs_lorenz(x, rho, sigma, beta)
    nodes = [dict(fn=lambda x: O.rk4_step(rhs, x, 0.01), inputs=["x"], outputs=["x"])]
    terms = [_term("obj", True, "eq", 2, 1.0, lambda d: d["x"], lambda d: 0.)]
    return dict(nodes=nodes, terms=terms, nsteps=nsteps, params=params, traj_keys=["x"])


# ---- gradients w.r.t. the white-box constants (system identification); `params` = the plant's
# nn.Parameters in registration order (reference ode.py), each a leaf the loss is differentiated by
def _t_sysid(params, nsteps, step, rhs, h, with_u, norm=2):
    if with_u:
        nodes = [dict(fn=lambda x, u: step(rhs, x, h, u), inputs=["xn", "U"], outputs=["xn"])]
    else:
        nodes = [dict(fn=lambda x: step(rhs, x, h), inputs=["xn"], outputs=["xn"])]
    terms = [_term("ref", True, "eq", norm, 1.0, lambda d: d["xn"][:, :-1, :], lambda d: d["X"])]
    return dict(nodes=nodes, terms=terms, nsteps=nsteps, params=params, traj_keys=["xn"])


def t_sysid_twotank(params, nsteps=8) -> dict:
    return _t_sysid(params, nsteps, O.rk4_step, lambda x, u: O.rhs_twotank(x, u, params[0], params[1]), 1.0, True)


def t_sysid_lorenz(params, nsteps=8) -> dict:
    return _t_sysid(params, nsteps, O.rk4_step, lambda x: O.rhs_lorenz(x, params[0], params[1], params[2]), 0.01, False)


def t_sysid_cstr(params, nsteps=6) -> dict:
    cc = dict(zip(("q", "V", "rho", "Cp", "mdelH", "EoverR", "k0", "UA", "Tf", "Caf"), params))
    return _t_sysid(params, nsteps, O.euler_step, lambda x, u: O.rhs_cstr(x, u, cc), 0.01, True)


def t_sysid_duffing(params, nsteps=9) -> dict:
    return _t_sysid(params, nsteps, O.rk2_step, lambda x, t: O.rhs_duffing(x, t, *params[:5]), 0.05, True)


def t_sysid_brusselator(params, nsteps=8) -> dict:
    return _t_sysid(params, nsteps, O.rk2_step, lambda x: O.rhs_brusselator(x, params[0], params[1]), 0.05, False, norm=1)


def t_sysid_lotka(params, nsteps=8) -> dict:
    return _t_sysid(params, nsteps, O.euler_step, lambda x: O.rhs_lotka_volterra(x, params[0]), 0.1, False)


def t_ude_brusselator(params, nsteps=6) -> dict:
    """universal differential equation (examples/ODEs/Part_3a_UDE.py:94-104): MLP 2-20-20-1 GELU inside the Brusselator,
    params = the MLP's W, b ... then alpha, beta"""
    block = lambda x: O.mlp_forward(params[:-2], x, "gelu")
    return _t_sysid(params, nsteps, O.rk4_step, lambda x: O.rhs_brusselator_hybrid(x, block, params[-2], params[-1]), 0.05, False)


def t_ude_lotka(params, nsteps=6) -> dict:
    """examples/ODEs/Part_3b_UDE.py: MLP inside Lotka-Volterra, params = the MLP's W, b ... then alpha, beta, delta, gamma"""
    block = lambda x: O.mlp_forward(params[:-4], x, "tanh")
    return _t_sysid(params, nsteps, O.euler_step, lambda x: O.rhs_lotka_volterra_hybrid(x, block, *params[-4:]), 0.1, False)


def t_sysid_vdp_policy(params, nsteps=10) -> dict:
    """closed loop: the policy (params[:-1]) and the plant's mu (params[-1]) are trained together"""
    rhs = lambda x, u: O.rhs_vdp(x, u, params[-1])
    nodes = [dict(fn=lambda x: O.mlp_forward(params[:-1], x, "tanh"), inputs=["x"], outputs=["u"]),
             dict(fn=lambda x, u: O.rk4_step(rhs, x, 0.1, u), inputs=["x", "u"], outputs=["x"])]
    terms = [_term("obj", True, "eq", 2, 1.0, lambda d: d["x"], lambda d: 0.)]
    return dict(nodes=nodes, terms=terms, nsteps=nsteps, params=params, traj_keys=["x", "u"])


def _t_preview(params, nsteps, pad_mode):
    """examples/control/Part_3_ref_tracking_ODE.py:93-105 shape: TwoTank + RK4 with a policy that previews the
    reference `r` (2 signals, window L+1 = 6) and a 1-d price signal `p` (window 3); `r` is SHORTER than
    nsteps + L so the window runs into the padding."""
    c1_, c2_ = _t(0.08, params[0]), _t(0.04, params[0])
    rhs = lambda x, u: O.rhs_twotank(x, u, c1_, c2_)
    nodes = [dict(name="policy", fn=lambda x, r, pz: O.bounds_scaling(O.mlp_forward(params, torch.cat([x, r, pz], dim=-1), "gelu"), 0., 1.),
                  inputs=["x", "r", "p"], outputs=["u"]),
             dict(name="model", fn=lambda x, u: O.rk4_step(rhs, x, 1.0, u), inputs=["x", "u"], outputs=["x"])]
    terms = [_term("track", True, "eq", 2, 5., lambda d: d["x"][:, 1:, :], lambda d: d["r"][:, :nsteps, :]),
             _term("lo", False, "gt", 1, 10., lambda d: d["x"], lambda d: 0.)]
    return dict(nodes=nodes, terms=terms, nsteps=nsteps, params=params, traj_keys=["x", "u"],
                preview=dict(keys_map={"r": ["policy"], "p": ["policy"]}, length={"r": 5, "p": 2}, pad_mode=pad_mode,
                             pad_constant=0.25 if pad_mode == "constant" else None))


def t_preview_replicate(params, nsteps=9): return _t_preview(params, nsteps, "replicate")
def t_preview_circular(params, nsteps=9): return _t_preview(params, nsteps, "circular")
def t_preview_constant(params, nsteps=9): return _t_preview(params, nsteps, "constant")
def t_preview_reflect(params, nsteps=9): return _t_preview(params, nsteps, "reflect")


CASES = {"c1": c1, "c2": c2, "c3": c3, "c4": c4, "c5": c5,
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
         "t_sysid_twotank": t_sysid_twotank, "t_sysid_lorenz": t_sysid_lorenz, "t_sysid_cstr": t_sysid_cstr,
         "t_sysid_duffing": t_sysid_duffing, "t_sysid_brusselator": t_sysid_brusselator, "t_sysid_lotka": t_sysid_lotka,
         "t_sysid_vdp_policy": t_sysid_vdp_policy, "t_ude_brusselator": t_ude_brusselator, "t_ude_lotka": t_ude_lotka,
         "t_barrier_log10": t_barrier_log10, "t_barrier_log": t_barrier_log, "t_barrier_inverse": t_barrier_inverse,
         "t_barrier_softexp": t_barrier_softexp, "t_barrier_softlog": t_barrier_softlog, "t_barrier_expshift": t_barrier_expshift}


def build(name: str, params: List[torch.Tensor], **kw) -> dict:
    return CASES[name](params, **kw)
