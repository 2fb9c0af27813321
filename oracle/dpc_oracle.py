"""
TEST INFRASTRUCTURE -- CPU oracle of the fused DPC rollout.  Not part of the product path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module; nothing under ``neuromancer_b200/`` does.

A plain-PyTorch (CPU, fp32 or fp64) *functional restatement* of the reference algorithm for the
hot path, op for op and in the reference's operation order, so that on CPU it is bit-identical to
running pnnl/neuromancer 1.5.6 itself (``oracle/pin_oracle.py`` asserts exactly that against the
unmodified reference wherever ``/root/reference`` exists, and writes the golden fixtures under
``tests/golden``).  Gradients come from torch autograd over the restated forward -- the same graph
the reference's ``loss.backward()`` (trainer.py:250) differentiates.

PARITY PIN: pinned against the real reference (torch.equal on trajectories, every loss term, the
total loss and all parameter gradients) for the 41 cases of ``oracle/cases.py`` (benchmark configurations, integrators,
white-box / hybrid right-hand sides, system identification, BarrierLoss, preview padding, structured maps) -- see
``tests/golden/PINNED.json`` written by ``oracle/pin_oracle.py``.  The reference's own test-suite
holds no value-level golden vectors for this path, so the fixtures generated
from the reference are the pin.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Sequence

import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------ blocks
def activation(name: str, param: float = 0.0) -> Callable[[torch.Tensor], torch.Tensor]:
    """torch.nn activations of the reference registry (modules/activations.py:230-256)."""
    table = {
        "identity": lambda z: z,
        "relu": F.relu,
        "gelu": F.gelu,                                     # nn.GELU(): exact erf form
        "leakyrelu": lambda z: F.leaky_relu(z, param if param else 0.01),
        "tanh": torch.tanh,
        "elu": lambda z: F.elu(z, param if param else 1.0),
        "sigmoid": torch.sigmoid,
        "softplus": F.softplus,
    }
    return table[name]


def mlp_forward(params: Sequence[torch.Tensor], x: torch.Tensor, act: str, bias: bool = True,
                act_param: float = 0.0) -> torch.Tensor:
    """``MLP.block_eval`` (reference blocks.py:169-177): ``x <- nlin(lin(x))`` per layer, the last
    nonlinearity is the identity.  ``params`` is the flat list W0, b0, W1, b1, ... (nn.Linear
    layout ``[out, in]``; slim/linear.py:104-119 -> ``F.linear``)."""
    fn = activation(act, act_param)
    step = 2 if bias else 1
    n_layers = len(params) // step
    for l in range(n_layers):
        w = params[step * l]
        b = params[step * l + 1] if bias else None
        x = F.linear(x, w, b)
        if l < n_layers - 1:
            x = fn(x)
    return x


def slim_perron_frobenius(weight, scaling, sigma_min, sigma_max):
    """``PerronFrobeniusLinear.effective_W`` (reference slim/linear.py:400-403): the ``[in, out]`` matrix"""
    s_clamped = sigma_max - (sigma_max - sigma_min) * torch.sigmoid(scaling)
    return s_clamped * F.softmax(weight, dim=1)


def slim_soft_svd(U, V, sigma, sigma_min, sigma_max):
    """``SVDLinear.effective_W`` (reference slim/linear.py:533-541): the ``[in, out]`` matrix"""
    s = sigma_max - (sigma_max - sigma_min) * torch.sigmoid(sigma)
    bounded = torch.eye(U.shape[0], V.shape[0], dtype=U.dtype) * s
    return torch.mm(U, torch.mm(bounded, V))


def bounds_scaling(x, xmin, xmax, scaling=1.0):
    """reference modules/functions.py:9-19"""
    return (xmax - xmin) * torch.sigmoid(scaling * x) + xmin


def bounds_clamp(x, xmin=None, xmax=None):
    """reference modules/functions.py:22-34"""
    if xmin is not None:
        x = x + torch.relu(-x + xmin)
    if xmax is not None:
        x = x - torch.relu(x - xmax)
    return x


# ------------------------------------------------------------------------------------------ dynamics
def rhs_twotank(x, u, c1, c2):
    """``TwoTankParam.ode_equations`` (reference dynamics/ode.py:227-237)"""
    h1 = torch.clip(x[:, [0]], min=0, max=1.0)
    h2 = torch.clip(x[:, [1]], min=0, max=1.0)
    pump = torch.clip(u[:, [0]], min=0, max=1.0)
    valve = torch.clip(u[:, [1]], min=0, max=1.0)
    dhdt1 = c1 * (1.0 - valve) * pump - c2 * torch.sqrt(h1)
    dhdt2 = c1 * valve * pump + c2 * torch.sqrt(h1) - c2 * torch.sqrt(h2)
    return torch.cat([dhdt1, dhdt2], dim=-1)


def rhs_vdp(x, u, mu):
    """``VanDerPolControl.ode_equations`` (reference dynamics/ode.py:363-368)"""
    x1 = x[:, [0]]
    x2 = x[:, [1]]
    dx1 = x2
    dx2 = mu * (1 - x1 ** 2) * x2 - x1 + u
    return torch.cat([dx1, dx2], dim=-1)


def rhs_cvdp(x, u, mu, kappa):
    """Ring of coupled ``VanDerPolControl`` oscillators (benchmark config C5):
    the reference equations per oscillator plus ``kappa*(x1[j+1] - x1[j])``."""
    pos, vel = x[:, 0::2], x[:, 1::2]
    nxt = torch.roll(pos, shifts=-1, dims=1)
    acc = mu * (1 - pos ** 2) * vel - pos + u + kappa * (nxt - pos)
    return torch.stack([vel, acc], dim=-1).reshape(x.shape[0], -1)


def rhs_lorenz_control(x, u, rho, sigma, beta):
    """``LorenzControl.ode_equations`` (reference dynamics/ode.py:404-414)"""
    x1, x2, x3 = x[:, [0]], x[:, [1]], x[:, [2]]
    u1, u2 = u[:, [0]], u[:, [1]]
    dx1 = sigma * (x2 - x1) + u1
    dx2 = x1 * (rho - x3) - x2
    dx3 = x1 * x2 - beta * x3 - u2
    return torch.cat([dx1, dx2, dx3], dim=-1)


def rhs_lorenz(x, rho, sigma, beta):
    """``LorenzParam.ode_equations`` (reference dynamics/ode.py:383-391)"""
    x1, x2, x3 = x[:, [0]], x[:, [1]], x[:, [-1]]
    return torch.cat([sigma * (x2 - x1), x1 * (rho - x3) - x2, x1 * x2 - beta * x3], dim=-1)


def rhs_cstr(x, u, c):
    """``CSTR_Param.ode_equations`` (reference dynamics/ode.py:449-461); ``c`` = dict of its constants"""
    Ca, T, Tc = x[:, [0]], x[:, [1]], u
    rA = c["k0"] * torch.exp(-c["EoverR"] / T) * Ca
    dCadt = c["q"] / c["V"] * (c["Caf"] - Ca) - rA
    dTdt = c["q"] / c["V"] * (c["Tf"] - T) + c["mdelH"] / (c["rho"] * c["Cp"]) * rA + c["UA"] / c["V"] / c["rho"] / c["Cp"] * (Tc - T)
    return torch.cat([dCadt, dTdt], dim=-1)


def rhs_duffing(x, u, alpha, beta, delta, gamma, omega):
    """``DuffingParam.ode_equations`` (reference dynamics/ode.py:254-262); the input is time"""
    x0, x1, t = x[:, [0]], x[:, [1]], u
    return torch.cat([x1, -delta * x1 - alpha * x0 - beta * x0 ** 3 + gamma * torch.cos(omega * t)], dim=-1)


def rhs_brusselator(x, alpha, beta):
    """``BrusselatorParam.ode_equations`` (reference dynamics/ode.py:276-281)"""
    x1, x2 = x[:, [0]], x[:, [-1]]
    return torch.cat([alpha + x2 * x1 ** 2 - beta * x1 - x1, beta * x1 - x2 * x1 ** 2], dim=-1)


def rhs_brusselator_hybrid(x, block, alpha, beta):
    """``BrusselatorHybrid.ode_equations`` (reference dynamics/ode.py:299-304); ``block`` = the unknown-term MLP 2 -> 1"""
    x1 = x[:, [0]]
    dx1 = alpha + block(x) - beta * x1 - x1
    dx2 = beta * x1 - block(x)
    return torch.cat([dx1, dx2], dim=-1)


def rhs_lotka_volterra_hybrid(x, block, alpha, beta, delta, gamma):
    """``LotkaVolterraHybrid.ode_equations`` (reference dynamics/ode.py:325-331)"""
    x1, x2 = x[:, [0]], x[:, [-1]]
    dx1 = alpha * x1 - beta * block(x)
    dx2 = delta * block(x) - gamma * x2
    return torch.cat([dx1, dx2], dim=-1)


def rhs_lotka_volterra(x, alpha):
    """``LotkaVolterraParam.ode_equations`` (reference dynamics/ode.py:344-349)"""
    x1, x2 = x[:, [0]], x[:, [-1]]
    return torch.cat([alpha * x1 - 0.4 * x1 * x2, 0.1 * x1 * x2 - 0.4 * x2], dim=-1)


def euler_step(f, x, h, *args):
    """``Euler.integrate`` (reference dynamics/integrators.py:153-156)"""
    k1 = f(x, *args)
    return x + h * k1


def rk2_step(f, x, h, *args):
    """``RK2.integrate`` (reference dynamics/integrators.py:189-193)"""
    k1 = f(x, *args)
    k2 = f(x + h * k1 / 2.0, *args)
    return x + h * k2


def rk4_step(f, x, h, *args):
    """``RK4.integrate`` (reference dynamics/integrators.py:205-211)"""
    k1 = f(x, *args)
    k2 = f(x + h * k1 / 2.0, *args)
    k3 = f(x + h * k2 / 2.0, *args)
    k4 = f(x + h * k3, *args)
    return x + h * (k1 / 6.0 + k2 / 3.0 + k3 / 3.0 + k4 / 6.0)


def euler_trap_step(f, x, h, *args):
    """``Euler_Trap.integrate`` (reference dynamics/integrators.py:169-177)"""
    pred = x + h * f(x, *args)
    corr = x + 0.5 * h * (f(x, *args) + f(pred, *args))
    return corr


def rk4_trap_step(f, x, h, *args):
    """``RK4_Trap.integrate`` (reference dynamics/integrators.py:229-235)"""
    k1 = f(x, *args)
    k2 = f(x + h * k1 / 2.0, *args)
    k3 = f(x + h * k2 / 2.0, *args)
    k4 = f(x + h * k3, *args)
    pred = x + h * (k1 / 6.0 + k2 / 3.0 + k3 / 3.0 + k4 / 6.0)
    corr = x + 0.5 * h * (f(x, *args) + f(pred, *args))
    return corr


def luther_step(f, x, h, *args):
    """``Luther.integrate`` (reference dynamics/integrators.py:247-264)"""
    q = 21 ** 0.5
    k1 = f(x, *args)
    k2 = f(x + h * k1, *args)
    k3 = f(x + h * (3 / 8 * k1 + 1 / 8 * k2), *args)
    k4 = f(x + h * (8 / 27 * k1 + 2 / 27 * k2 + 8 / 27 * k3), *args)
    k5 = f(x + h * ((-21 + 9 * q) / 392 * k1 + (-56 + 8 * q) / 392 * k2 + (336 - 48 * q) / 392 * k3 + (-63 + 3 * q) / 392 * k4), *args)
    k6 = f(x + h * ((-1155 - 255 * q) / 1960 * k1 + (-280 - 40 * q) / 1960 * k2 - 320 * q / 1960 * k3 + (63 + 363 * q) / 1960 * k4
                    + (2352 + 392 * q) / 1960 * k5), *args)
    k7 = f(x + h * ((330 + 105 * q) / 180 * k1 + 120 / 180 * k2 + (-200 + 280 * q) / 180 * k3 + (126 - 189 * q) / 180 * k4
                    + (-686 - 126 * q) / 180 * k5 + (490 - 70 * q) / 180 * k6), *args)
    return x + h * (1 / 20 * k1 + 16 / 45 * k3 + 49 / 180 * k5 + 49 / 180 * k6 + 1 / 20 * k7)


def rkf_step(f, x, h, *args):
    """``Runge_Kutta_Fehlberg.integrate`` (reference dynamics/integrators.py:287-301), 5th-order output"""
    k1 = f(x, *args)
    k2 = f(x + h * k1 / 4, *args)
    k3 = f(x + 3 * h * k1 / 32 + 9 * h * k2 / 32, *args)
    k4 = f(x + h * k1 * 1932 / 2197 - 7200 / 2197 * h * k2 + 7296 / 2197 * h * k3, *args)
    k5 = f(x + h * k1 * 439 / 216 - 8 * h * k2 + 3680 / 513 * h * k3 - 845 / 4104 * h * k4, *args)
    k6 = f(x - 8 / 27 * h * k1 + 2 * h * k2 - 3544 / 2565 * h * k3 + 1859 / 4104 * h * k4 - 11 / 40 * h * k5, *args)
    return x + h * (k1 * 16 / 135 + k3 * 6656 / 12825 + k4 * 28561 / 56430 - 9 / 50 * k5 + k6 * 2 / 55)


def leapfrog_step(f, X, h, *args):
    """``LeapFrog.integrate`` (reference dynamics/integrators.py:350-360): ddx = f(x) on X = [x, dx]"""
    n = X.shape[-1] // 2
    x, dx = X[:, :n], X[:, n:2 * n]
    x_1 = x + dx * h + 0.5 * f(x, *args) * h ** 2
    ddx_1 = f(x_1, *args)
    dx_1 = dx + 0.5 * (f(x, *args) + ddx_1) * h
    return torch.cat([x_1, dx_1], dim=-1)


def yoshida4_step(f, X, h, *args):
    """``Yoshida4.integrate`` (reference dynamics/integrators.py:373-404)"""
    n = X.shape[-1] // 2
    x, dx = X[:, :n], X[:, n:2 * n]
    w0 = -2 ** (1 / 3) / (2 - 2 ** (1 / 3))
    w1 = 1 / (2 - 2 ** (1 / 3))
    c1 = w1 / 2
    c4 = c1
    c2 = (w0 + w1) / 2
    c3 = c2
    d1 = w1
    d3 = d1
    d2 = w0
    x_1 = x + c1 * dx * h
    dx_1 = dx + d1 * f(x_1, *args) * h
    x_2 = x_1 + c2 * dx_1 * h
    dx_2 = dx_1 + d2 * f(x_2, *args) * h
    x_3 = x_2 + c3 * dx_2 * h
    dx_3 = dx_2 + d3 * f(x_3, *args) * h
    x_4 = x_3 + c4 * dx_3 * h
    return torch.cat([x_4, dx_3], dim=-1)


INTEGRATORS = {"leapfrog": leapfrog_step, "yoshida4": yoshida4_step, "euler": euler_step, "rk2": rk2_step, "rk4": rk4_step, "euler_trap": euler_trap_step,
               "rk4_trap": rk4_trap_step, "luther": luther_step, "rkf": rkf_step}


def ssm_step(x, u, d, A, B, E=None):
    """``SSM.forward`` with bias-free linear maps (reference dynamics/ode.py:33-51):
    ``x = fx(x) + fu(u)`` then ``x += fd(d)``."""
    out = F.linear(x, A) + F.linear(u, B)
    if E is not None and d is not None:
        out = out + F.linear(d, E)
    return out


# ------------------------------------------------------------------------------------------ rollout
def system_forward(nodes: List[dict], data: Dict[str, torch.Tensor], nsteps: int) -> Dict[str, torch.Tensor]:
    """``System.forward`` (reference system.py:261-277) with ``Node.forward`` (:48-59) and
    ``System.cat`` (:234-246): for every step and node, read index ``i`` of each input key, call the
    node, append the outputs on the time axis.  ``nodes`` = ``[{fn, inputs, outputs}, ...]``."""
    data = dict(data)
    for i in range(nsteps):
        for node in nodes:
            args = [data[k][:, i] for k in node["inputs"]]
            out = node["fn"](*args)
            if not isinstance(out, tuple):
                out = (out,)
            for k, v in zip(node["outputs"], out):
                if k not in data:
                    data[k] = v[:, None, :]
                else:
                    data[k] = torch.cat([data[k], v[:, None, :]], dim=1)
    return data


def preview_window(sig: torch.Tensor, i: int, length: int, pad_mode: str, pad_constant=None) -> torch.Tensor:
    """``SystemPreview.get_data_with_preview`` (reference system.py:325-345): the window
    ``sig[:, i : i+1+length, :]``; if it is short, the FULL signal is right-padded by ``length`` samples
    with ``torch.nn.functional.pad(mode=pad_mode)`` and the window is taken from that."""
    data = sig[:, i:i + 1 + length, :]
    if data.shape[1] < length + 1:
        data = F.pad(sig, (0, 0, 0, length), mode=pad_mode, value=pad_constant)[:, i:i + 1 + length, :]
    return data


def system_preview_forward(nodes: List[dict], data: Dict[str, torch.Tensor], nsteps: int, preview_keys_map: dict,
                           preview_length: dict, pad_mode: str, pad_constant=None) -> Dict[str, torch.Tensor]:
    """``SystemPreview.forward`` (reference system.py:348-367).  ``nodes`` entries carry a ``name``."""
    data = dict(data)
    for i in range(nsteps):
        for node in nodes:
            args = []
            for k in node["inputs"]:
                if k in preview_keys_map and node["name"] in preview_keys_map[k]:
                    win = preview_window(data[k], i, preview_length[k], pad_mode, pad_constant)
                    args.append(win.reshape(data[k].size(0), -1))
                else:
                    args.append(data[k][:, i])
            out = node["fn"](*args)
            if not isinstance(out, tuple):
                out = (out,)
            for k, v in zip(node["outputs"], out):
                if k not in data:
                    data[k] = v[:, None, :]
                else:
                    data[k] = torch.cat([data[k], v[:, None, :]], dim=1)
    return data


# ------------------------------------------------------------------------------------------ loss
def compare(cmp: str, norm: int, left: torch.Tensor, right: torch.Tensor):
    """``LT/GT/Eq.forward`` (reference constraint.py:85-98, 122-134, 157-173) -> (loss, value, violation)"""
    if cmp == "lt":
        value = left - right
        penalty = F.relu(value)
        if norm == 2:
            penalty = penalty ** 2
        return torch.mean(penalty), value, penalty
    if cmp == "gt":
        value = right - left
        penalty = F.relu(value)
        if norm == 2:
            penalty = penalty ** 2
        return torch.mean(penalty), value, penalty
    value = left - right
    left_b, right_b = torch.broadcast_tensors(left, right)
    if norm == 1:
        return F.l1_loss(left_b, right_b), value, torch.abs(value)
    return F.mse_loss(left_b, right_b), value, value ** 2


def penalty_loss(terms: List[dict], data: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """``PenaltyLoss.forward`` (reference loss.py:168-181) over ``Constraint.forward``
    (constraint.py:307-324).  ``terms`` = ``[{name, objective, cmp, norm, weight, left, right}]`` with
    ``left`` / ``right`` callables ``data -> tensor | float`` (the ``Variable`` expression).
    Sums start from the python float 0.0 and run in list order: objectives, then constraints."""
    out = {}
    obj, pen = 0.0, 0.0
    for t in [t for t in terms if t["objective"]] + [t for t in terms if not t["objective"]]:
        left, right = t["left"](data), t["right"](data)
        # Constraint.__init__ wraps python numbers with torch.tensor(.) (constraint.py:245-262); the
        # fp64 attribution runs of the tests follow the dtype of the other side instead of float32
        like = left if isinstance(left, torch.Tensor) else (right if isinstance(right, torch.Tensor) else None)
        dt = like.dtype if like is not None else torch.float32
        left = left if isinstance(left, torch.Tensor) else torch.tensor(left, dtype=dt)
        right = right if isinstance(right, torch.Tensor) else torch.tensor(right, dtype=dt)
        loss, value, violation = compare(t["cmp"], t["norm"], left, right)
        out[t["name"]] = t["weight"] * loss
        out[t["name"] + "_value"] = value
        out[t["name"] + "_violation"] = violation
        if t["objective"]:
            obj = obj + out[t["name"]]
        else:
            pen = pen + out[t["name"]]
    out["objective_loss"], out["penalty_loss"] = obj, pen
    out["loss"] = obj + pen
    return out


def barrier_function(name: str, alpha: float, shift: float) -> Callable[[torch.Tensor], torch.Tensor]:
    """the built-in barrier functions of ``BarrierLoss.__init__`` (reference loss.py:209-215)"""
    return {"log10": lambda value: -torch.log10(-value),
            "log": lambda value: -torch.log(-value),
            "inverse": lambda value: 1 / (-value),
            "softexp": lambda value: (torch.exp(alpha * value) - 1) / alpha + alpha,
            "softlog": lambda value: -torch.log(1 + alpha * (-value - alpha)) / alpha,
            "expshift": lambda value: torch.exp(value + shift)}[name]


def barrier_loss(terms: List[dict], data: Dict[str, torch.Tensor], barrier: str, upper_bound: float = 1., shift: float = 1.,
                 alpha: float = 0.5, out_of_place: bool = False) -> Dict[str, torch.Tensor]:
    """``BarrierLoss`` (reference loss.py:184-257): ``PenaltyLoss.forward`` with ``calculate_constraints`` replaced --
    entries with value >= 0 pay their violation, the others the barrier value (nan / inf -> 0, clamped to
    [0, upper_bound]); each part a masked mean over ALL entries, the per-constraint keys keep the plain penalty.
    ``out_of_place``: the reference's masked assignments modify the barrier function's output in place; for 'expshift'
    that output is what ExpBackward saved and the reference's own ``loss.backward()`` raises.  Same values, with the
    assignments on a copy, define the gradient the fused kernels are checked against for that barrier."""
    out = penalty_loss(terms, data)
    fn = barrier_function(barrier, alpha, shift)
    loss, b_loss = 0.0, 0.0
    for t in [t for t in terms if not t["objective"]]:
        cvalue, cviolation = out[t["name"] + "_value"], out[t["name"] + "_violation"]
        penalty_mask = cvalue >= 0
        cbarrier = fn(cvalue).clone() if out_of_place else fn(cvalue)
        cbarrier[cbarrier != cbarrier] = 0.0
        cbarrier[cbarrier == float("Inf")] = 0.0
        cbarrier = torch.clamp(cbarrier, min=0.0, max=upper_bound)
        out[t["name"] + "_barrier"] = cbarrier
        if penalty_mask.any():
            loss = loss + t["weight"] * torch.mean(penalty_mask * cviolation)
        if (~penalty_mask).any():
            bl = t["weight"] * torch.mean(~penalty_mask * cbarrier)
            b_loss = b_loss + bl
            loss = loss + bl
    out["barrier_loss"], out["penalty_loss"] = b_loss, loss
    out["loss"] = out["objective_loss"] + loss
    return out


def run_case(spec: dict, data: Dict[str, torch.Tensor], want_grad: bool = True) -> Dict[str, object]:
    """Forward (+ backward) of one problem.  ``spec`` = {nodes, terms, nsteps, params (list of leaf
    tensors), traj_keys}.  Returns trajectories, per-term scalars, loss and parameter gradients."""
    if "preview" in spec:
        pv = spec["preview"]
        traj = system_preview_forward(spec["nodes"], data, spec["nsteps"], pv["keys_map"], pv["length"], pv["pad_mode"],
                                      pv.get("pad_constant"))
    else:
        traj = system_forward(spec["nodes"], data, spec["nsteps"])
    scored = barrier_loss(spec["terms"], traj, **spec["barrier"]) if "barrier" in spec else penalty_loss(spec["terms"], traj)
    res = {"traj": {k: traj[k] for k in spec["traj_keys"]},
           "terms": {t["name"]: scored[t["name"]] for t in spec["terms"]},
           "objective_loss": scored["objective_loss"], "penalty_loss": scored["penalty_loss"],
           "loss": scored["loss"]}
    if want_grad:
        leaves = [p for p in spec["params"] if p.requires_grad]
        grads = torch.autograd.grad(scored["loss"], leaves, allow_unused=True) if leaves else []
        res["grads"] = [g if g is not None else torch.zeros_like(p) for g, p in zip(grads, leaves)]
    return res
