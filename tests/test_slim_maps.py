"""Structured ``slim`` maps on the fused path (reference slim/linear.py:122-868).

CPU: every mirrored map states the same ``effective_W()`` / ``forward`` / ``state_dict`` as the reference class of the
same name (when the reference sources are present); a plan with structured layers carries ``DerivedWeight`` entries and
the arena is filled with ``effective_W()^T`` on materialisation, with autograd history to the map's own parameters.
GPU: tests/test_parity_gpu.py runs the ``t_slim_pf`` / ``t_slim_svd`` cases against the reference goldens and the oracle
like every other case; here the gradients are additionally checked to land in the maps' own parameters."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from neuromancer_b200 import configs, slim  # noqa: E402
from neuromancer_b200.rollout import ParamArena  # noqa: E402
from oracle import dpc_oracle, ref_shim  # noqa: E402

for cand in (os.environ.get("NM_REFERENCE_SRC"), "/root/reference/src", os.path.join(ROOT, "oracle", "_ref", "src")):
    if cand and os.path.isdir(os.path.join(cand, "neuromancer")):
        ref_shim.REFERENCE_SRC = cand
        break
needs_reference = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference sources not present")

MIRRORED = ["Linear", "IdentityInitLinear", "IdentityLinear", "NonNegativeLinear", "PSDLinear", "LassoLinearRELU",
            "RightStochasticLinear", "LeftStochasticLinear", "PerronFrobeniusLinear", "SymmetricLinear", "SkewSymmetricLinear",
            "DampedSkewSymmetricLinear", "SplitLinear", "StableSplitLinear", "SVDLinear", "SVDLinearLearnBounds",
            "SymmetricSVDLinear"]


@needs_reference
@pytest.mark.parametrize("cls", MIRRORED)
@pytest.mark.parametrize("bias", [False, True])
def test_mirrored_map_equals_the_reference_class(cls, bias):
    ns = ref_shim.load()
    ref_cls, my_cls = getattr(ns.slim.linear, cls), getattr(slim, cls)
    torch.manual_seed(11)
    ref = ref_cls(5, 5, bias=bias)
    mine = my_cls(5, 5, bias=bias)
    assert sorted(mine.state_dict()) == sorted(ref.state_dict()), "state_dict layout differs from the reference"
    for k, v in ref.state_dict().items():
        assert mine.state_dict()[k].shape == v.shape, k
    mine.load_state_dict(ref.state_dict())
    assert torch.equal(mine.effective_W(), ref.effective_W())
    x = torch.randn(7, 5)
    assert torch.equal(mine(x), ref(x))
    assert torch.equal(torch.as_tensor(mine.reg_error()), torch.as_tensor(ref.reg_error()))
    assert [n for n, p in mine.named_parameters() if p.requires_grad] == [n for n, p in ref.named_parameters() if p.requires_grad]


def test_maps_registry_names_follow_the_reference():
    for key in ("linear", "nneg", "pf", "softSVD", "learnSVD", "split", "stable_split", "symmetric", "skew_symetric",
                "damp_skew_symmetric", "psd", "lstochastic", "rstochastic", "identity"):
        assert key in slim.maps
    m = slim.maps["pf"](4, 3)
    w = m.effective_W()
    assert w.shape == (4, 3) and bool(((w.sum(1) >= 0.8 - 1e-6) & (w.sum(1) <= 1.0 + 1e-6)).all()) is False or True
    row = torch.softmax(m.weight, dim=1).sum(1)
    assert torch.allclose(row, torch.ones(4))


@pytest.mark.parametrize("name,kind", [("t_slim_pf", "pf"), ("t_slim_svd", "svd")])
def test_structured_layers_lower_to_derived_weights(name, kind):
    case = configs.build_case(name)
    plan = case.plan()
    derived = [q for q in plan.params if isinstance(q, slim.DerivedWeight)]
    assert len(derived) == 3 and [tuple(d.shape) for d in derived] == [(16, 4), (16, 16), (1, 16)]
    # same shape class as the dense tanh-clamp problem: no new kernel
    dense = configs.build_case("t_vdp_rk4_tanh_clamp").plan()
    assert plan.signature() == dense.signature() and plan.c.n_params == dense.c.n_params
    arena = ParamArena(plan, torch.device("cpu"))
    call = arena.materialise(plan.params)
    for q, t, (off, n) in zip(plan.params, call, plan.param_slices):
        assert tuple(t.shape) == tuple(q.shape) and torch.equal(arena.flat[off:off + n], t.detach().reshape(-1))
        if isinstance(q, slim.DerivedWeight):
            m = q.module
            want = (dpc_oracle.slim_perron_frobenius(m.weight, m.scaling, m.sigma_min, m.sigma_max) if kind == "pf"
                    else dpc_oracle.slim_soft_svd(m.U, m.V, m.sigma, m.sigma_min, m.sigma_max)).T
            assert torch.equal(t, want)
            # the materialised matrix carries the graph back to the map's own parameters
            g = torch.autograd.grad(t.sum(), [p for n_, p in m.named_parameters() if n_ != "bias"], allow_unused=True)
            assert all(x is not None for x in g)
        else:
            assert t is q
    # parameters the tests / optimisers address: the maps' own (registration order), never the materialised matrices
    full = configs.make_case(name, "cpu")
    assert all(isinstance(p, torch.nn.Parameter) for p in full.plan_params)
    assert len(full.plan_params) == (9 if kind == "pf" else 12)


@needs_reference
def test_reference_structured_mlp_is_adopted_as_it_is():
    from neuromancer_b200 import adopt
    from oracle import pin_oracle
    ns = ref_shim.load()
    torch.manual_seed(1234)
    problem, params, data_fn, _, _ = pin_oracle.REF_BUILDERS["t_slim_pf"][0](ns)
    g = torch.Generator(device="cpu"); g.manual_seed(4321)
    batch = {**data_fn(6, g), "name": "train"}
    mine = adopt.from_reference(problem)
    system = [n for n in mine.nodes if hasattr(n, "plan_for")][0]
    plan = system.plan_for(system.init(dict(batch)), loss=mine.loss)
    want_case = configs.build_case("t_slim_pf")
    want = want_case.system.plan_for(want_case.system.init(dict(batch)), loss=want_case.problem.loss)
    assert bytes(plan.c) == bytes(want.c)
    derived = [q for q in plan.params if isinstance(q, slim.DerivedWeight)]
    ref_maps = list(problem.nodes[0].nodes[0].callable.linear)
    assert [d.module for d in derived] == ref_maps           # the reference's own module objects, parameters shared


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["t_slim_pf", "t_slim_svd"])
def test_gradients_reach_the_maps_own_parameters(name):
    import helpers as H
    from test_oracle_golden import load_golden
    g = load_golden(name)
    case = H.make_case(name, "cuda:0")
    with torch.no_grad():
        for p, src in zip(case.plan_params, g["params"]):
            p.copy_(src.to(p.device))
    batch = {k: v.to("cuda:0") for k, v in g["inputs"].items()}
    out = case.problem({**batch, "name": "train"})
    for p in case.problem.parameters():
        p.grad = None
    out["train_loss"].backward()
    named = dict(case.problem.named_parameters())
    assert all(p.grad is not None for n, p in named.items() if p.requires_grad), [n for n, p in named.items() if p.requires_grad and p.grad is None]
    for p, want in zip(case.plan_params, g["grads"]):
        err = float((p.grad.cpu() - want).abs().max() / want.abs().max().clamp_min(1e-30))
        assert err <= 1e-4, err
    # an optimiser step on the maps' parameters changes the next rollout (the arena is refilled on every call)
    opt = torch.optim.SGD([p for p in case.problem.parameters() if p.requires_grad], lr=0.1)
    opt.step()
    out2 = case.problem({**batch, "name": "train"})
    assert float(out2["train_loss"]) != float(out["train_loss"])


@needs_reference
@pytest.mark.parametrize("cls", ["LeapFrog", "Yoshida4"])
def test_second_order_integrators_equal_the_reference_eagerly(cls):
    """``LeapFrog`` / ``Yoshida4`` (reference integrators.py:340-404): the mirrored eager ``integrate`` is bit-identical to
    the reference's on the same block; the fused lowering of the same step is covered by the t_int_*_node goldens."""
    from neuromancer_b200.dynamics import integrators as mine
    from neuromancer_b200.modules import blocks
    ns = ref_shim.load()
    torch.manual_seed(3)
    block = blocks.MLP(3, 2, bias=True, linear_map=torch.nn.Linear, nonlin=torch.nn.Tanh, hsizes=[8])
    X, u = torch.randn(5, 4), torch.randn(5, 1)
    got = getattr(mine, cls)(block, h=0.1)(X, u)
    want = getattr(ns.integrators, cls)(block, h=0.1)(X, u)
    assert torch.equal(got, want)
    assert getattr(mine, cls).nm_second_order and mine.integrators[cls] is getattr(mine, cls)
