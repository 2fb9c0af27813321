"""API-mirror behaviour on CPU, following the reference's own unit tests for the classes on the path:
tests/test_constraints.py (operator identities, zero / non-zero violations, norms, slicing),
tests/test_loss.py (linearity of PenaltyLoss under + and *), tests/test_functions.py (bounds),
tests/test_blocks.py / test_integrators.py / test_ode.py (shapes), tests/test_system.py (Node keys)."""
import pytest
import torch
import torch.nn as nn

from neuromancer_b200 import Node, PenaltyLoss, variable
from neuromancer_b200.constraint import Constraint, Variable
from neuromancer_b200.dynamics import integrators, ode
from neuromancer_b200.modules import blocks
from neuromancer_b200.modules.functions import bounds_clamp, bounds_scaling


def test_variable_arithmetic_matches_tensor_math():
    d = {"x": torch.randn(5, 4, 3), "y": torch.randn(5, 4, 3)}
    x, y = variable("x"), variable("y")
    for expr, want in [(x + y, d["x"] + d["y"]), (x - 2.0, d["x"] - 2.0), (3 * x * y, 3 * d["x"] * d["y"]),
                       (x / 2, d["x"] / 2), (-x, -d["x"]), (x ** 2, d["x"] ** 2), (abs(x), d["x"].abs()),
                       (x[:, 1:, :] - x[:, :-1, :], d["x"][:, 1:, :] - d["x"][:, :-1, :]),
                       (x[:, [-1], :], d["x"][:, [-1], :]), (torch.sin(x), torch.sin(d["x"])),
                       (x @ y.mT, d["x"] @ d["y"].mT)]:
        assert torch.equal(expr(d), want)
    assert set((x + y).keys) == {"x", "y"}


@pytest.mark.parametrize("norm", [1, 2])
def test_comparators_zero_when_satisfied_nonzero_when_violated(norm):
    x = variable("x")
    lo, hi = torch.zeros(6, 2), torch.ones(6, 2)
    inside = {"x": torch.full((6, 2), 0.5)}
    outside = {"x": torch.full((6, 2), 2.0)}
    for con, ok, bad in [((x < hi) ^ norm, inside, outside), ((x > lo) ^ norm, inside, {"x": -outside["x"]}),
                         ((x == 0.5) ^ norm, inside, outside)]:
        assert float(con(ok)[con.key]) == 0.0
        assert float(con(bad)[con.key]) != 0.0
    c = (x < hi) ^ 2
    val = c(outside)
    assert torch.equal(val[f"{c.key}_value"], outside["x"] - hi)
    assert torch.equal(val[f"{c.key}_violation"], (outside["x"] - hi) ** 2)


def test_constraint_weight_norm_and_keys():
    x, r = variable("x"), variable("r")
    c = 5.0 * ((x == r) ^ 2)
    assert isinstance(c, Constraint) and c.weight == 5.0 and c.comparator.norm == 2
    assert c.key == "x_eq_r" and c.output_keys == ["x_eq_r", "x_eq_r_value", "x_eq_r_violation"]
    c.update_name("track")
    assert c.output_keys[0] == "track"
    d = {"x": torch.randn(3, 4, 2), "r": torch.randn(3, 4, 2)}
    assert torch.isclose(c(d)["track"], 5.0 * torch.nn.functional.mse_loss(d["x"], d["r"]))


def test_penalty_loss_sum_and_linearity():
    x = variable("x")
    o1, o2 = (x == 1.0) ^ 2, 0.5 * (x == 0.0)
    c1 = 3.0 * (x < 0.2)
    for t, n in ((o1, "o1"), (o2, "o2"), (c1, "c1")):
        t.name = n
    loss = PenaltyLoss([o1, o2], [c1])
    d = {"x": torch.rand(7, 3, 2, dtype=torch.float64)}
    out = loss(d)
    assert torch.isclose(out["loss"], out["objective_loss"] + out["penalty_loss"])
    assert out["C_values"].shape == (7, 6) and out["C_eq_values"].shape == (7, 0) and out["C_ineq_violations"].shape == (7, 6)
    doubled = (2.0 * loss)(d)
    assert torch.isclose(doubled["loss"], 2.0 * out["loss"])
    other = PenaltyLoss([(x == 2.0) ^ 2], [1.5 * (x > 0.9)])
    summed = (loss + other)(d)
    assert torch.isclose(summed["loss"], out["loss"] + other(d)["loss"])


def test_bounds_stay_inside_interval():
    x = 10 * torch.randn(500, 5)
    lo, hi = -torch.ones(500, 5), torch.ones(500, 5)
    for fn in (bounds_scaling, bounds_clamp):
        y = fn(x, lo, hi)
        assert torch.all(y <= hi) and torch.all(y >= lo)


def test_block_integrator_ode_shapes():
    net = blocks.MLP_bounds(insize=5, outsize=2, hsizes=[8, 8], nonlin=nn.ELU, min=-1.0, max=1.0, method="relu_clamp")
    assert net(torch.randn(9, 3), torch.randn(9, 2)).shape == (9, 2)       # Block.forward concatenates inputs
    for cls, nx, nu in ((ode.TwoTankParam, 2, 2), (ode.VanDerPolControl, 2, 1)):
        f = cls()
        for integ in (integrators.Euler, integrators.RK2, integrators.RK4):
            assert integ(f, h=0.1)(torch.rand(4, nx), torch.rand(4, nu)).shape == (4, nx)
    rhs = blocks.MLP(3, 3, hsizes=[6], nonlin=nn.Tanh)
    assert integrators.RK4(rhs, h=0.05)(torch.randn(4, 3)).shape == (4, 3)
    ring = ode.CoupledVanDerPolControl(n_osc=6)
    assert ring(torch.randn(4, 12), torch.randn(4, 6)).shape == (4, 12)
    ssm = ode.SSM(nn.Linear(3, 3, bias=False), nn.Linear(2, 3, bias=False), 3, 2, fd=nn.Linear(1, 3, bias=False), nd=1)
    assert ssm(torch.randn(4, 3), torch.randn(4, 2), torch.randn(4, 1)).shape == (4, 3)


def test_node_key_handling():
    n = Node(lambda a, b: (a + b, a - b), [variable("p"), "q"], ["s", variable("d")], name="n")
    assert n.input_keys == ["p", "q"] and n.output_keys == ["s", "d"]
    out = n({"p": torch.ones(2, 1), "q": torch.ones(2, 1), "unused": None})
    assert set(out) == {"s", "d"} and float(out["s"][0]) == 2.0
    single = Node(lambda a: 2 * a, ["p"], ["o"])
    assert list(single({"p": torch.ones(1, 1)})) == ["o"]
    net = blocks.MLP(1, 1, hsizes=[2], nonlin=nn.ReLU)
    node = Node(net, ["p"], ["o"])
    node.freeze()
    assert not any(p.requires_grad for p in net.parameters())
    node.unfreeze()
    assert all(p.requires_grad for p in net.parameters())


def test_device_prefetcher_keeps_order_and_content_on_cpu():
    import torch
    from neuromancer_b200.dataset import DevicePrefetcher
    batches = [{"x": torch.full((2, 3), float(i)), "name": "train"} for i in range(5)]
    got = list(DevicePrefetcher(iter(batches), device="cpu", depth=2))
    assert len(got) == 5
    for i, b in enumerate(got):
        assert b["name"] == "train" and torch.equal(b["x"], batches[i]["x"])
    assert list(DevicePrefetcher(iter([]), device="cpu")) == []


@pytest.mark.parametrize("barrier", ["log10", "log", "inverse", "softexp", "softlog", "expshift"])
def test_barrier_loss_matches_reference_golden(barrier):
    """BarrierLoss (reference loss.py:184-257) against vectors recorded from the unmodified reference
    (oracle/pin_barrier.py): same torch ops in the same order, so equality is exact."""
    import os
    import numpy as np
    from neuromancer_b200 import constraint as nb_constraint, loss as nb_loss
    from oracle.pin_barrier import build
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "barrier.npz"))
    data = {"x": torch.from_numpy(g["in_x"]), "u": torch.from_numpy(g["in_u"])}
    res = build(nb_constraint, nb_loss, barrier)(data)
    for k in ("loss", "objective_loss", "penalty_loss", "barrier_loss", "x_hi_barrier", "x_lo_barrier"):
        want = torch.from_numpy(g[f"{barrier}/{k}"])
        got = torch.as_tensor(res[k]).detach()
        assert got.shape == want.shape and torch.equal(got, want), (barrier, k, got, want)
    assert float(res["loss"]) == float(res["objective_loss"] + res["penalty_loss"])


def test_problem_fuses_penalty_loss_and_named_barrier_loss():
    from neuromancer_b200 import BarrierLoss, Problem, System
    from neuromancer_b200.dynamics import integrators, ode
    from neuromancer_b200.modules import blocks
    net = blocks.MLP(2, 1, hsizes=[8], nonlin=nn.ReLU)
    system = System([Node(net, ["x"], ["u"], name="policy"),
                     Node(integrators.Euler(ode.VanDerPolControl(), h=0.1), ["x", "u"], ["x"], name="model")], nsteps=4)
    x = variable("x")
    con = (x < 1.0)
    assert Problem([system], PenaltyLoss([], [con]))._fusable()
    # BarrierLoss with a named barrier is lowered (barrier kind in the term descriptors); a callable barrier, or any other
    # subclass that redefines how constraints enter the loss, is evaluated with torch ops on the fused trajectories
    assert Problem([system], BarrierLoss([], [con], barrier="softexp"))._fusable()
    assert not Problem([system], BarrierLoss([], [con], barrier=lambda v: torch.exp(v)))._fusable()

    class MyLoss(PenaltyLoss):
        pass
    assert not Problem([system], MyLoss([], [con]))._fusable()
