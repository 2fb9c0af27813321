"""Host-side logic (no GPU): node-list recognition, NmPlan lowering, loss-term lowering, error
behaviour of the drop-in boundary.  Mirrors what the reference's tests/test_system.py and
tests/test_problem.py check about key wiring, with the fused path's stricter contract."""
import pytest
import torch
import torch.nn as nn

from neuromancer_b200 import Node, PenaltyLoss, Problem, System, _lib as L, configs, variable
from neuromancer_b200.dynamics import integrators, ode
from neuromancer_b200.modules import blocks
from neuromancer_b200.plan import PlanError, lower_term, recognise


def test_c2_plan_fields():
    case = configs.build_case("c2")
    plan = case.plan()
    p = plan.c
    assert (p.nx, p.nu, p.ny, p.nsteps) == (2, 2, 0, 30)
    assert p.has_policy == 1 and [p.policy.sizes[i] for i in range(4)] == [4, 32, 32, 2]
    assert p.policy.act == 2 and p.policy.bounds == L.NM_BOUNDS_SIGMOID_SCALE
    assert (p.policy.bmin[0], p.policy.bmax[1]) == (0.0, 1.0)
    assert (p.dyn_kind, p.integrator, p.rhs_kind) == (L.NM_DYN_ODE, L.NM_INT_RK4, L.NM_RHS_TWOTANK)
    assert abs(p.rhs_const[0] - 0.08) < 1e-7 and abs(p.rhs_const[1] - 0.04) < 1e-7 and p.h == 1.0
    assert p.n_segs == 2 and p.segs[0].kind == L.NM_SEG_STATE and p.segs[1].kind == L.NM_SEG_EXO
    assert p.n_exo == 1 and (p.exo[0].width, p.exo[0].steps) == (2, 31)
    assert p.n_params == 4 * 32 + 32 + 32 * 32 + 32 + 32 * 2 + 2 == 1282
    assert p.n_terms == 5 and not plan.python_terms
    # terminal constraint x[:, [-1], :] > r - 0.01 : fixed time entry 30 broadcast against 31 steps of r
    t = p.terms[3]
    assert (t.cmp, t.norm, t.t_len, t.d_len) == (L.NM_CMP_GT, 1, 31, 2) and abs(t.weight - 10.0) < 1e-6
    assert (t.lhs.a.var, t.lhs.a.t_start, t.lhs.a.t_len) == (L.NM_VAR_X, 30, 1)
    assert t.rhs.a.var == L.NM_VAR_EXO0 and abs(t.rhs.a.offset + 0.01) < 1e-9 and t.rhs.a.scale == 1.0
    assert [bool(p.terms[i].is_objective) for i in range(5)] == [True, False, False, False, False]


def test_c3_finite_difference_terms_and_loss_only_input():
    plan = configs.build_case("c3").plan()
    p = plan.c
    assert p.has_policy == 0 and p.rhs_kind == L.NM_RHS_MLP and p.nsteps == 100
    assert plan.exo_keys == ["X"]                      # only the loss reads the measured trajectory
    fd = p.terms[1]
    assert fd.lhs.op == L.NM_OP_SUB and fd.rhs.op == L.NM_OP_SUB and fd.t_len == 99
    assert (fd.lhs.a.var, fd.lhs.a.t_start, fd.lhs.b.t_start) == (L.NM_VAR_EXO0, 1, 0)
    assert (fd.rhs.a.var, fd.rhs.a.t_start, fd.rhs.b.t_start, fd.rhs.a.t_len) == (L.NM_VAR_X, 1, 0, 99)


def test_c4_preview_segments_and_product_term():
    plan = configs.build_case("c4").plan()
    p = plan.c
    assert (p.nx, p.nu, p.ny) == (4, 1, 1) and p.has_output == 1 and p.dyn_kind == L.NM_DYN_SSM
    assert [p.segs[i].width for i in range(p.n_segs)] == [1, 48, 48, 48, 48]
    assert p.dist_exo == 4 and p.exo[4].width == 3
    price = p.terms[0]
    assert price.lhs.op == L.NM_OP_MUL and price.lhs.a.var == L.NM_VAR_U
    assert (price.lhs.b.t_len, price.lhs.b.d_start, price.lhs.b.d_len) == (96, 0, 1)
    assert price.rhs.a.var == L.NM_VAR_CONST and price.rhs.a.offset == 0.0
    assert abs(p.C[3] - 1.0) < 1e-7 and abs(p.A[0] - 0.95) < 1e-7


def test_unrecognised_callable_raises_not_implemented():
    net = blocks.MLP(2, 1, hsizes=[4], nonlin=nn.ReLU)
    lam = Node(lambda x, u: x + u, ["x", "u"], ["x"], name="plant")
    system = System([Node(net, ["x"], ["u"], name="pol"), lam], nsteps=3)
    with pytest.raises(NotImplementedError, match="not recognised"):
        system.layout()


def test_misordered_nodes_and_missing_keys():
    plant = ode.VanDerPolControl()
    plant.mu.requires_grad_(False)
    dyn = Node(integrators.RK4(plant, h=0.1), ["x", "u"], ["x"], name="model")
    pol = Node(blocks.MLP(2, 1, hsizes=[4], nonlin=nn.ReLU), ["x"], ["u"], name="pol")
    with pytest.raises(PlanError, match="order"):
        System([dyn, pol], nsteps=3).layout()
    good = System([pol, dyn], nsteps=3)
    with pytest.raises(KeyError):                           # reference: missing data key -> KeyError
        good.plan_for({"r": torch.zeros(2, 4, 2)})
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        good({"x": torch.zeros(2, 1, 2)})


def test_trainable_white_box_constants_become_a_derived_arena_entry():
    """system identification (reference ode.py:357-361: mu is a trainable nn.Parameter): the constant vector is lowered as a
    derived entry behind the policy's parameters, flagged through rhs_mlp.trainable / .param_offset; frozen constants are
    plain plan values tracked by version"""
    from neuromancer_b200.dynamics.ode import DerivedConstants
    plant = ode.VanDerPolControl()                          # mu requires_grad=True by default
    dyn = Node(integrators.RK4(plant, h=0.1), ["x", "u"], ["x"], name="model")
    pol = Node(blocks.MLP(2, 1, hsizes=[4], nonlin=nn.ReLU), ["x"], ["u"], name="pol")
    system = System([pol, dyn], nsteps=3)
    plan = system.plan_for({"x": torch.zeros(2, 1, 2)})
    n_pol = sum(q.numel() for q in pol.callable.parameters())
    assert isinstance(plan.params[-1], DerivedConstants) and plan.param_slices[-1] == (n_pol, 1)
    assert plan.c.n_params == n_pol + 1 and plan.c.rhs_mlp.trainable == 1 and plan.c.rhs_mlp.param_offset == n_pol
    w = plan.params[-1].materialise()
    assert w.requires_grad and torch.autograd.grad(w.sum(), plant.mu)[0].item() == 1.0
    with torch.no_grad():
        plant.mu.add_(1.0)                                  # an optimiser step does not make the plan stale ...
    assert not plan.values_stale()
    plant.mu.requires_grad_(False)                          # ... frozen constants are carried by value and tracked
    plan2 = System([pol, dyn], nsteps=3).plan_for({"x": torch.zeros(2, 1, 2)})
    assert plan2.c.n_params == n_pol and plan2.c.rhs_mlp.trainable == 0 and plan2.c.rhs_const[0] == 2.0
    with torch.no_grad():
        plant.mu.add_(1.0)
    assert plan2.values_stale()


def test_unsupported_activation_and_structured_map():
    from neuromancer_b200.modules.activations import SoftExponential
    pol = Node(blocks.MLP(2, 1, hsizes=[4]), ["x"], ["u"], name="pol")   # default nonlin = SoftExponential
    assert isinstance(pol.callable.nonlin[0], SoftExponential)
    plant = ode.VanDerPolControl()
    plant.mu.requires_grad_(False)
    dyn = Node(integrators.Euler(plant, h=0.1), ["x", "u"], ["x"], name="model")
    with pytest.raises(NotImplementedError, match="SoftExponential"):
        System([pol, dyn], nsteps=2).plan_for({"x": torch.zeros(2, 1, 2)})


def test_term_lowering_rules():
    ids = {"x": L.NM_VAR_X, "u": L.NM_VAR_U, "r": L.NM_VAR_EXO0}
    shapes = {"x": (11, 3), "u": (10, 2), "r": (11, 3)}
    x, u, r = variable("x"), variable("u"), variable("r")
    t = lower_term(2.0 * ((x[:, 1:, 0:2] - 0.5) < r[:, :-1, [1]]) ^ 2, False, ids, shapes)
    assert t is not None and (t.t_len, t.d_len, t.norm, t.cmp) == (10, 2, 2, L.NM_CMP_LT)
    assert (t.lhs.a.t_start, t.lhs.a.d_len, t.lhs.a.offset) == (1, 2, -0.5)
    assert (t.rhs.a.d_start, t.rhs.a.d_len, t.rhs.a.t_len) == (1, 1, 10)
    assert lower_term((-x) == 3 * r, True, ids, shapes).lhs.a.scale == -1.0
    assert lower_term(torch.tanh(x) == 0.0, True, ids, shapes) is None        # opaque torch function
    assert lower_term(x[:, ::2, :] == 0.0, True, ids, shapes) is None         # strided slice
    assert lower_term(x == u, True, ids, shapes) is None                      # shapes do not broadcast
    assert lower_term(variable("z") == 0.0, True, ids, shapes) is None        # unknown key
    w = torch.tensor(2.0, requires_grad=True)
    assert lower_term(w * (x == 0.0), True, ids, shapes) is None              # trainable weight


def test_problem_output_keys_are_lazy_and_prefixed():
    from neuromancer_b200.problem import LazyOutputs
    lz = LazyOutputs({"a": 1})
    calls = []
    lz.register(["b", "c"], lambda: calls.append(1) or {"b": 2, "c": 3})
    assert "b" in lz and "zzz" not in lz and not calls
    assert lz["c"] == 3 and lz["b"] == 2 and len(calls) == 1
    with pytest.raises(KeyError):
        lz["zzz"]


def test_param_order_matches_reference_state_dict_names():
    net = blocks.MLP_bounds(insize=4, outsize=2, hsizes=[32, 32], nonlin=nn.GELU)
    keys = list(net.state_dict().keys())
    assert keys[:4] == ["linear.0.linear.weight", "linear.0.linear.bias", "linear.1.linear.weight", "linear.1.linear.bias"] \
        or "linear.0.weight" in keys
    for k in ("linear.0.weight", "linear.0.bias", "linear.0.linear.weight", "linear.2.linear.bias"):
        assert k in keys, keys
    assert net.layer_sizes() == [4, 32, 32, 2]


def test_recognise_layouts():
    lay = recognise(list(configs.build_case("c4").system.nodes))
    assert (lay.state_key, lay.action_key, lay.output_key, lay.dist_key) == ("x", "u", "y", "d")
    assert lay.exo_keys[:4] == ["d_obs_lookahead", "ymin_lookahead", "ymax_lookahead", "price_lookahead"]


def test_constant_division_of_the_rk4_combine_is_correctly_rounded():
    """nm_dynamics.cuh::div_const replaces x / 6 and x / 3 by  q0 = RN(x r), q = RN(q0 + (x - d q0) r)  with r = RN(1 / d):
    identical to the IEEE quotient on EVERY significand of a binade (division by a constant commutes with the exponent)
    and on random values of other magnitudes -- emulated here with exact double arithmetic (a product of two floats and
    the residual are exact in double)."""
    import numpy as np
    rng = np.random.default_rng(0)
    binade = np.arange(1 << 23, 1 << 24, dtype=np.int64).astype(np.float32)
    spread = np.concatenate([rng.standard_normal(1 << 20).astype(np.float32) * np.float32(s) for s in (1.0, 1e-3, 1e3, 1e-20, 1e20)])
    for d in (3.0, 6.0):
        r = np.float32(1.0 / d)
        for x in (binade, -binade, spread):
            q0 = (x.astype(np.float64) * np.float64(r)).astype(np.float32)
            rem = x.astype(np.float64) - np.float64(d) * q0.astype(np.float64)
            assert np.array_equal(rem.astype(np.float32).astype(np.float64), rem)          # the FMA residual is exact
            q = (q0.astype(np.float64) + rem * np.float64(r)).astype(np.float32)
            assert np.array_equal(q, (x / np.float32(d)).astype(np.float32))


def test_hybrid_right_hand_side_lowering():
    """grey-box right-hand sides (reference ode.py:284-331): the block is lowered as the right-hand-side MLP, the constants as a
    derived entry behind it; rhs_mlp.trainable bit 0 = weights, bit 1 = constants"""
    from neuromancer_b200.dynamics.ode import DerivedConstants
    net = blocks.MLP(2, 1, bias=True, linear_map=nn.Linear, nonlin=nn.GELU, hsizes=[8])
    plant = ode.BrusselatorHybrid(net)
    system = System([Node(integrators.RK4(plant, h=0.05), ["x"], ["x"], name="ude")], nsteps=4)
    plan = system.plan_for({"x": torch.zeros(3, 1, 2)})
    p = plan.c
    n_mlp = sum(q.numel() for q in net.parameters())
    assert p.rhs_kind == L.NM_RHS_BRUSS_HYBRID and L.rhs_has_mlp(p.rhs_kind) and p.has_policy == 0
    assert [p.rhs_mlp.sizes[i] for i in range(p.rhs_mlp.n_layers + 1)] == [2, 8, 1]
    assert p.rhs_mlp.trainable == 3 and p.n_params == n_mlp + 2
    assert isinstance(plan.params[-1], DerivedConstants) and plan.param_slices[-1] == (n_mlp, 2)
    assert [q is r for q, r in zip(plan.params[-1].own_parameters(), (plant.alpha, plant.beta))] == [True, True]
    assert (p.rhs_const[0], p.rhs_const[1]) == (5.0, 5.0)
    for q in (plant.alpha, plant.beta):
        q.requires_grad_(False)
    frozen = System([Node(integrators.RK4(plant, h=0.05), ["x"], ["x"], name="ude")], nsteps=4).plan_for({"x": torch.zeros(3, 1, 2)})
    assert frozen.c.rhs_mlp.trainable == 1 and frozen.c.n_params == n_mlp
    for q in net.parameters():
        q.requires_grad_(False)
    plant.alpha.requires_grad_(True)
    only_theta = System([Node(integrators.RK4(plant, h=0.05), ["x"], ["x"], name="ude")], nsteps=4).plan_for({"x": torch.zeros(3, 1, 2)})
    assert only_theta.c.rhs_mlp.trainable == 2 and only_theta.c.n_params == n_mlp + 2
    with pytest.raises(PlanError, match="first-order"):
        System([Node(integrators.LeapFrog(plant, h=0.05), ["x"], ["x"], name="ude")], nsteps=4).plan_for({"x": torch.zeros(3, 1, 2)})
    with pytest.raises(AssertionError):
        ode.LotkaVolterraHybrid(blocks.MLP(3, 1, hsizes=[4], nonlin=nn.Tanh))          # the block maps x (2) -> 1


def test_barrier_loss_lowering():
    """BarrierLoss with a named barrier (reference loss.py:184-257): kind in NmTerm.norm bits 4-7 of the CONSTRAINTS only,
    parameters in the plan; a change of a parameter makes the cached plan stale; a callable barrier is not lowered"""
    from neuromancer_b200 import BarrierLoss
    case = configs.build_case("t_barrier_softlog")
    plan = case.plan()
    p = plan.c
    assert plan.barrier == L.NM_BARRIERS["softlog"] == 5 and not plan.python_terms
    assert [p.terms[k].norm for k in range(p.n_terms)] == [2, 2 | 5 << 4, 1 | 5 << 4, 1 | 5 << 4]
    assert [bool(p.terms[k].is_objective) for k in range(p.n_terms)] == [True, False, False, False]
    assert (round(p.barrier_alpha, 6), round(p.barrier_shift, 6), p.barrier_ub) == (0.4, 0.7, 2.0)
    assert not plan.values_stale()
    case.problem.loss.upper_bound = 3.0
    assert plan.values_stale()
    loss = case.problem.loss
    eager = BarrierLoss(list(loss.objectives), list(loss.constraints), barrier=lambda v: torch.exp(v))
    assert eager.barrier_name is None
    batch = case.make_batch(2, "cpu", 0)
    unfused = case.system.plan_for(case.system.init(dict(batch)), loss=eager)
    assert unfused.barrier == 0 and all(unfused.c.terms[k].norm in (1, 2) for k in range(unfused.c.n_terms))
