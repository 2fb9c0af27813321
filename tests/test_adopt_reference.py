"""Drop-in boundary from the REFERENCE side: problems built with the unmodified
reference classes (loaded through oracle/ref_shim.py: build container, or the oracle/_ref copy on the GPU box) are
adopted by ``neuromancer_b200.adopt.from_reference`` and must lower to the very same POD ``NmPlan`` as the same
problem written against this package -- byte for byte -- while sharing the reference model's Parameter objects."""
import ctypes as C
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from neuromancer_b200 import adopt, configs  # noqa: E402
from neuromancer_b200.dynamics import ode  # noqa: E402
from oracle import ref_shim  # noqa: E402

for cand in (os.environ.get("NM_REFERENCE_SRC"), "/root/reference/src", os.path.join(ROOT, "oracle", "_ref", "src")):
    if cand and os.path.isdir(os.path.join(cand, "neuromancer")):
        ref_shim.REFERENCE_SRC = cand
        break
pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference sources not present")

NAMES = ["c1", "c2", "c3", "c4", "c5"]


def _frozen(ns, w):
    lin = torch.nn.Linear(w.shape[1], w.shape[0], bias=False)
    with torch.no_grad():
        lin.weight.copy_(w)
    lin.weight.requires_grad_(False)
    return lin


def _with_modules(ns, name, problem):
    """c1 / c4 are written with python lambdas in the examples (and in oracle/pin_oracle.py); a lambda is opaque, so the
    adoptable form of the same systems uses the reference's own ode.SSM / nn.Linear modules"""
    from oracle import cases as ocases
    system = problem.nodes[0]
    if name == "c1":
        A, B = torch.tensor([[1.2, 1.0], [0.0, 1.0]]), torch.tensor([[1.0], [0.5]])
        system.nodes[1].callable = ns.ode.SSM(_frozen(ns, A), _frozen(ns, B), 2, 1)
    else:
        A, Bm, E, Cm = (torch.tensor(m) for m in (ocases.C4_A, ocases.C4_B, ocases.C4_E, ocases.C4_C))
        system.nodes[2].callable = ns.ode.SSM(_frozen(ns, A), _frozen(ns, Bm), 4, 1, fd=_frozen(ns, E), nd=3)
        system.nodes[0].callable = _frozen(ns, Cm)
    return problem


def _build(name, modules=True):
    from oracle import pin_oracle
    ns = ref_shim.load()
    torch.manual_seed(1234)
    problem, params, data_fn, traj_keys, term_names = pin_oracle.REF_BUILDERS[name][0](ns)
    if name in ("c1", "c4") and modules:
        problem = _with_modules(ns, name, problem)
    g = torch.Generator(device="cpu"); g.manual_seed(4321)
    return problem, params, {**data_fn(6, g), "name": "train"}


@pytest.mark.parametrize("name", ["c1", "c4"])
def test_python_lambdas_are_refused_loudly(name):
    from neuromancer_b200.plan import PlanError
    problem, _, _ = _build(name, modules=False)
    with pytest.raises(PlanError, match="python function is opaque"):
        adopt.from_reference(problem)


# the C5 plant of oracle/pin_oracle.py is a user subclass of the reference's ODESystem, not a reference class
adopt.register_rhs("CoupledVdP", lambda ref: ode.CoupledVanDerPolControl(n_osc=6, mu=float(ref.mu), kappa=float(ref.kappa)))


@pytest.mark.parametrize("name", NAMES)
def test_reference_built_problem_lowers_to_the_same_plan(name):
    ref_problem, ref_params, batch = _build(name)
    assert type(ref_problem).__module__.startswith("neuromancer.")
    mine = adopt.from_reference(ref_problem)
    system = [n for n in mine.nodes if hasattr(n, "plan_for")][0]
    plan = system.plan_for(system.init(dict(batch)), loss=mine.loss)
    want_case = configs.build_case(name)
    want = want_case.system.plan_for(want_case.system.init(dict(batch)), loss=want_case.problem.loss)
    a, b = bytes(plan.c), bytes(want.c)
    if a != b:                                                      # name the first differing field
        for fname, _ in type(plan.c)._fields_:
            fa, fb = getattr(plan.c, fname), getattr(want.c, fname)
            ba = bytes(fa) if hasattr(fa, "_length_") or hasattr(fa, "_fields_") else fa
            bb = bytes(fb) if hasattr(fb, "_length_") or hasattr(fb, "_fields_") else fb
            assert ba == bb, f"{name}: NmPlan.{fname} differs between the adopted reference model and the native one"
    assert a == b
    # every python-evaluated (not lowered) term would show up here: all terms of c1..c5 are fused
    assert not plan.python_terms and plan.c.n_terms == want.c.n_terms
    # the adopted model trains the reference model: same Parameter objects, same flat order
    assert len(plan.params) == len(ref_params)
    for p, q in zip(plan.params, ref_params):
        assert p is q


@pytest.mark.parametrize("name", ["t_sysid_twotank", "t_sysid_cstr", "t_sysid_vdp_policy", "t_ude_brusselator", "t_ude_lotka"])
def test_reference_built_identification_problem_lowers_to_the_same_plan(name):
    """system identification / universal differential equations built with the reference's ode classes (trainable
    nn.Parameter constants, ode.py:216-237, 284-331, 417-461): same NmPlan bytes as the native build, and the derived
    constant entry differentiates the reference model's own Parameter objects"""
    ref_problem, ref_params, batch = _build(name)
    mine = adopt.from_reference(ref_problem)
    system = [n for n in mine.nodes if hasattr(n, "plan_for")][0]
    plan = system.plan_for(system.init(dict(batch)), loss=mine.loss)
    want_case = configs.build_case(name)
    with torch.no_grad():                                           # same values: the plan carries the constants by value
        flat = []
        for q in want_case.plan().params:
            flat += q.own_parameters() if hasattr(q, "materialise") else [q]
        for dst, src in zip(flat, ref_params):
            dst.copy_(src)
    want = want_case.system.plan_for(want_case.system.init(dict(batch)), loss=want_case.problem.loss)
    assert bytes(plan.c) == bytes(want.c)
    leaves = []
    for q in plan.params:
        leaves += q.own_parameters() if hasattr(q, "materialise") else [q]
    assert len(leaves) == len(ref_params) and all(a is b for a, b in zip(leaves, ref_params))
    assert plan.c.rhs_mlp.trainable in (1, 3)


def test_adopted_constraint_values_match_the_reference_on_cpu():
    """the adopted loss terms evaluate (eagerly, torch CPU) to what the reference's own Constraint objects give"""
    ref_problem, _, batch = _build("c2")
    mine = adopt.from_reference(ref_problem)
    x = torch.rand(6, 31, 2)
    u = torch.rand(6, 30, 2)
    data = {"x": x, "u": u, "r": batch["r"]}
    ref_out = ref_problem.loss({**data, "name": "train"}) if False else None   # reference PenaltyLoss needs all keys incl. name
    ref_terms = [t for t in list(ref_problem.loss.objectives) + list(ref_problem.loss.constraints)]
    my_terms = mine.loss.terms()
    assert len(ref_terms) == len(my_terms)
    for rt, mt in zip(ref_terms, my_terms):
        want = rt(data)[rt.output_keys[0]]
        got = mt(data)[mt.output_keys[0]]
        assert torch.equal(torch.as_tensor(got), torch.as_tensor(want)), (rt.name, got, want)
    assert ref_out is None


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c2", "c5", "t_slim_pf", "t_sysid_cstr", "t_ude_brusselator"])
def test_reference_built_problem_runs_on_the_fused_kernels(name):
    """End of the drop-in chain on a GPU box (oracle/_ref travels with the snapshot): a problem built with the reference's
    classes, adopted, run through nmfused::rollout_fwd / rollout_bwd, against the golden recorded from the SAME reference
    problem on the CPU -- and the gradients land in the reference model's own Parameter objects."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_oracle_golden import load_golden
    from test_parity_gpu import compare
    g = load_golden(name)
    ref_problem, ref_params, _ = _build(name)
    with torch.no_grad():
        for p, src in zip(ref_params, g["params"]):
            p.copy_(src)
    mine = adopt.from_reference(ref_problem).to("cuda:0")
    batch = {k: v.to("cuda:0") for k, v in g["inputs"].items()}
    out = mine({**batch, "name": "train"})
    for p in ref_params:
        p.grad = None
    out["train_loss"].backward()
    names = [str(n) for n in g["terms"]]
    got = {"traj": {k: out[f"train_{k}"].detach().cpu() for k in g["traj"]},
           "terms": {n: out[f"train_{n}"].detach().cpu() for n in names},
           "objective_loss": out["train_objective_loss"].detach().cpu(), "penalty_loss": out["train_penalty_loss"].detach().cpu(),
           "loss": out["train_loss"].detach().cpu(), "grads": [p.grad.detach().cpu() for p in ref_params]}
    compare(got, g, f"{name}: adopted reference-built problem vs reference golden")
