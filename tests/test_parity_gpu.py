"""GPU parity tests proper: the CUDA product path (through the public API -> C ABI -> sm_100a kernels)
against (a) the golden vectors recorded from the unmodified reference and (b) the CPU oracle on fresh
seeded inputs, incl. ragged batch sizes.  Tolerances are the north_star's: trajectories rtol 1e-5,
loss and gradients rtol 1e-4 (relative to the tensor's scale)."""
import ctypes as C

import pytest
import torch

import helpers as H
from test_oracle_golden import load_golden

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
NAMES = sorted(H.configs.CASES)
TRAJ_RTOL, GRAD_RTOL = 1e-5, 1e-4


def assert_rel(got, want, rtol, what):
    err = H.rel_err(got, want)
    assert err <= rtol, f"{what}: max |diff| / max |ref| = {err:.3e} > {rtol:g}"


def to_dev(batch):
    return {k: (v.to(DEV) if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}


def load_params(case, params):
    with torch.no_grad():
        for p, src in zip(case.plan_params, params):
            p.copy_(src.to(p.device))


def compare(got, ref, tag, grad_rtol=None):
    grad_rtol = GRAD_RTOL if grad_rtol is None else grad_rtol
    for k, v in ref["traj"].items():
        assert got["traj"][k].shape == v.shape
        assert_rel(got["traj"][k], v, TRAJ_RTOL, f"{tag} trajectory {k}")
    loss_scale = float(torch.as_tensor(ref["loss"]).detach().abs())
    for k, v in ref["terms"].items():
        # a penalty term that is ~0 next to the total loss (a handful of barely-violated bounds) is
        # judged on the loss scale: its relative error is ill-conditioned by construction
        v = torch.as_tensor(v).detach()
        err = float((got["terms"][k].double() - v.double()).abs())
        assert err <= GRAD_RTOL * max(float(v.abs()), 1e-3 * loss_scale), f"{tag} term {k}: |diff| = {err:.3e}"
    for k in ("objective_loss", "penalty_loss", "loss"):
        assert_rel(got[k], torch.as_tensor(ref[k]).detach(), GRAD_RTOL, f"{tag} {k}")
    assert len(got["grads"]) == len(ref["grads"])
    for i, (a, b) in enumerate(zip(got["grads"], ref["grads"])):
        assert_rel(a, b, grad_rtol, f"{tag} grad {i}")


@pytest.mark.parametrize("name", NAMES)
def test_matches_reference_golden(name):
    g = load_golden(name)
    case = H.make_case(name, DEV)
    load_params(case, g["params"])
    batch = to_dev({**g["inputs"], "name": "train"})
    got = H.product_run(case, batch)
    compare(got, g, f"{name} vs reference golden")


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("B", [1, 127, 129, 300])
def test_matches_oracle_ragged_batches(name, B):
    if name == "c5" and B > 129:
        B = 160
    case = H.make_case(name, DEV, seed=3)
    batch_cpu = case.make_batch(B, "cpu", 7)
    params_cpu = [p.detach().cpu().clone() for p in case.plan_params]
    got = H.product_run(case, to_dev(batch_cpu))
    ref = H.oracle_run(name, params_cpu, batch_cpu)
    compare(got, ref, f"{name} B={B} vs oracle")


def test_native_library_is_loaded_and_counts_launches():
    from neuromancer_b200 import _lib
    lib = _lib.load()
    case = H.make_case("c2", DEV)
    out = case.problem(to_dev(case.make_batch(64, "cpu", 0)))
    assert lib.nm_last_launch_count() == 3      # weight prep, rollout (+fused scoring), loss finalise
    out["train_loss"].backward()
    assert lib.nm_last_launch_count() == 3      # weight prep, adjoint (+fused term gradients), gradient finalise
    maps = open("/proc/self/maps").read()
    assert "libnm_rollout.so" in maps


def test_deterministic_and_independent_of_batch_padding():
    """Same inputs -> bitwise identical trajectories; a trajectory's rollout does not depend on which
    other trajectories share its tile."""
    case = H.make_case("c2", DEV)
    b = to_dev(case.make_batch(257, "cpu", 5))
    a1 = case.problem(b)["train_x"].clone()
    a2 = case.problem(b)["train_x"].clone()
    assert torch.equal(a1, a2)
    sub = {k: (v[100:131].contiguous() if isinstance(v, torch.Tensor) else v) for k, v in b.items()}
    a3 = case.problem(sub)["train_x"]
    assert torch.equal(a3, a1[100:131])


def test_system_forward_alone_and_no_grad():
    case = H.make_case("t_vdp_rk4_tanh_clamp", DEV)
    batch = to_dev(case.make_batch(50, "cpu", 2))
    with torch.no_grad():
        out = case.system({k: v for k, v in batch.items() if k != "name"})
    assert out["x"].shape == (50, case.nsteps + 1, 2) and out["u"].shape == (50, case.nsteps, 1)
    assert not out["x"].requires_grad
    full = case.problem(batch)
    assert torch.equal(full["train_x"], out["x"])


def test_gradient_is_linear_in_upstream_scale():
    case = H.make_case("c2", DEV)
    batch = to_dev(case.make_batch(200, "cpu", 9))
    params = [p for p in case.plan_params if p.requires_grad]

    def grads(scale):
        for p in params:
            p.grad = None
        (scale * case.problem(batch)["train_loss"]).backward()
        return [p.grad.clone() for p in params]
    g1, g3 = grads(1.0), grads(3.0)
    for a, b in zip(g1, g3):      # not bitwise: split-group partial sums meet through atomics
        assert H.rel_err(3.0 * a, b) < 1e-5


def test_backward_through_single_term_and_lazy_outputs():
    """d(term_k) alone goes through grad_scalars[k]; value/violation tensors materialise lazily and
    agree with the reference formula."""
    name = "t_vdp_rk4_tanh_clamp"
    case = H.make_case(name, DEV)
    batch_cpu = case.make_batch(64, "cpu", 4)
    out = case.problem(to_dev(batch_cpu))
    params = [p for p in case.plan_params if p.requires_grad]
    box = [t for t in case.problem.loss.terms() if t.name == "box"][0]
    key = f"train_{box.output_keys[0]}"
    gbox = torch.autograd.grad(out[key], params, retain_graph=True, allow_unused=True)
    pc = [p.detach().cpu().clone().requires_grad_(True) for p in case.plan_params]
    spec = H.ocases.build(name, pc)
    traj = H.dpc_oracle.system_forward(spec["nodes"], batch_cpu, spec["nsteps"])
    scored = H.dpc_oracle.penalty_loss(spec["terms"], traj)
    want = torch.autograd.grad(scored["box"], pc, allow_unused=True)
    for a, b in zip(gbox, want):
        if b is None:
            continue
        assert_rel(a.cpu(), b, GRAD_RTOL, "single-term gradient")
    val = out[f"train_{box.output_keys[1]}"]
    assert_rel(val.cpu(), scored["box_value"].detach(), TRAJ_RTOL, "lazy constraint value")
    assert out["train_C_values"].shape == (64, 13 * 2)
    assert out["train_C_eq_values"].shape == (64, 0)


def test_python_evaluated_term_feeds_adjoint_kernel():
    """A term the plan compiler cannot lower (here: a torch function of the trajectory) is evaluated
    with torch ops and its gradient enters the adjoint kernel through grad_x_ext."""
    from neuromancer_b200 import PenaltyLoss, Problem, variable
    case = H.make_case("t_vdp_euler_elu", DEV)
    x = variable("x")
    smooth = 0.3 * (torch.tanh(x) == 0.1) ^ 2
    smooth.name = "smooth"
    fused = (variable("u") == 0.) ^ 2
    fused.name = "fused"
    problem = Problem([case.system], PenaltyLoss([smooth, fused], [])).to(DEV)
    batch_cpu = case.make_batch(90, "cpu", 6)
    out = problem(to_dev(batch_cpu))
    plan = case.system.plan_for(case.system.init(to_dev(batch_cpu)), loss=problem.loss)
    assert len(plan.python_terms) == 1 and plan.c.n_terms == 1
    params = [p for p in case.plan_params if p.requires_grad]
    got = torch.autograd.grad(out["train_loss"], params)
    pc = [p.detach().cpu().clone().requires_grad_(True) for p in case.plan_params]
    spec = H.ocases.build("t_vdp_euler_elu", pc)
    traj = H.dpc_oracle.system_forward(spec["nodes"], batch_cpu, spec["nsteps"])
    loss = 0.3 * torch.nn.functional.mse_loss(torch.tanh(traj["x"]), torch.full_like(traj["x"], 0.1)) \
        + torch.mean(traj["u"] ** 2)
    want = torch.autograd.grad(loss, pc)
    assert_rel(out["train_loss"].detach().cpu(), loss.detach(), GRAD_RTOL, "loss with python term")
    for a, b in zip(got, want):
        assert_rel(a.cpu(), b, GRAD_RTOL, "gradient with python term")


def test_mask_conventions_bit_exact():
    """ReLU / clamp / relu-penalty masks at exactly representable kinks follow torch autograd:
    relu'(0)=0, clip passes gradient on the closed interval, sign(0)=0."""
    from neuromancer_b200 import Node, PenaltyLoss, Problem, System, variable
    from neuromancer_b200.dynamics import integrators, ode
    from neuromancer_b200.modules import blocks
    plant = ode.VanDerPolControl()
    plant.mu = torch.nn.Parameter(torch.tensor(0.7), requires_grad=False)
    net = blocks.MLP(insize=2, outsize=1, hsizes=[8], nonlin=torch.nn.ELU, bias=False)
    system = System([Node(net, ["x"], ["u"], name="policy"),
                     Node(integrators.Euler(plant, h=0.05), ["x", "u"], ["x"], name="model")], nsteps=9)
    x, u = variable("x"), variable("u")
    t1 = (x[:, [0], :] > 1.0)            # value = 1 - x0 : exactly 0 where x0 == 1 -> relu'(0) = 0
    t2 = (x[:, [0], :] == 2.0)           # norm-1 equality: sign(0) = 0 where x0 == 2
    t3 = (u == 0.) ^ 2
    for t_, n_ in ((t1, "t1"), (t2, "t2"), (t3, "t3")):
        t_.name = n_
    problem = Problem([system], PenaltyLoss([t3], [t1, t2])).to(DEV)
    x0 = torch.tensor([[[1.0, 2.0]], [[0.5, 2.5]], [[2.0, -1.0]]], device=DEV)
    x0.requires_grad_(True)
    out = problem({"x": x0, "name": "train"})
    (out[f"train_{t1.output_keys[0]}"] + out[f"train_{t2.output_keys[0]}"]).backward()
    xr = x0.detach().cpu().clone().requires_grad_(True)
    v1 = torch.relu(1.0 - xr[:, [0], :]).mean()
    v2 = torch.nn.functional.l1_loss(xr[:, [0], :], torch.full_like(xr[:, [0], :], 2.0))
    (v1 + v2).backward()
    assert torch.equal(x0.grad.cpu(), xr.grad), (x0.grad, xr.grad)


@pytest.mark.parametrize("pad_mode", ["replicate", "circular", "constant", "reflect"])
def test_preview_window_indexing_bit_exact(pad_mode):
    """SystemPreview (reference system.py:325-367) is pure indexing: with a one-hot linear policy and
    x+ = u the state trajectory must equal, bit for bit, the entries torch.nn.functional.pad + slicing
    select -- including windows that run past the end of the signal."""
    import torch.nn.functional as F
    case = H.make_case("t_preview_index", DEV, pad_mode=pad_mode)
    T, L, d = case.nsteps, 3, 2
    batch_cpu = case.make_batch(37, "cpu", 1)
    net = case.system.nodes[0].callable
    for pick in [(0, 1), (5, 2), (7, 6)]:                     # which flattened window entry feeds u_0 / u_1
        with torch.no_grad():
            w = torch.zeros(2, (L + 1) * d)
            w[0, pick[0]] = 1.0
            w[1, pick[1]] = 1.0
            net.linear[0].weight.copy_(w.to(DEV))
        with torch.no_grad():
            out = case.system({k: v.to(DEV) for k, v in batch_cpu.items() if k != "name"})
        sig = batch_cpu["s"]
        padded = F.pad(sig, (0, 0, 0, L), mode=pad_mode, value=-7.5 if pad_mode == "constant" else None)
        want = torch.zeros(37, T + 1, 2)
        for i in range(T):
            win = sig[:, i:i + 1 + L, :]
            if win.shape[1] < L + 1:
                win = padded[:, i:i + 1 + L, :]
            flat = win.reshape(37, -1)
            want[:, i + 1, 0] = flat[:, pick[0]]
            want[:, i + 1, 1] = flat[:, pick[1]]
        assert torch.equal(out["x"].cpu(), want), (pad_mode, pick)


def test_c4_preview_formulation_equals_materialised_lookahead():
    """C4 with SystemPreview windows over the raw signals reproduces C4 with materialised look-ahead
    tensors (grid_response.ipynb).  Since round 2 the materialised formulation runs with its exogenous first-layer
    part hoisted into a GEMM in front of the rollout (tensor-core policy mode), the windowed one on the FFMA kernels:
    different summation orders, so the two agree to the trajectory / gradient tolerances instead of bitwise."""
    a = H.make_case("c4", DEV, seed=2)
    b = H.make_case("c4p", DEV, seed=2)
    with torch.no_grad():
        for q, src in zip(b.plan_params, a.plan_params):
            q.copy_(src)
    ra = H.product_run(a, to_dev(a.make_batch(150, "cpu", 3)))
    rb = H.product_run(b, to_dev(b.make_batch(150, "cpu", 3)))
    for k in ("x", "u", "y"):
        assert_rel(rb["traj"][k], ra["traj"][k], TRAJ_RTOL, f"trajectory {k}")
    assert_rel(rb["loss"], ra["loss"], 1e-5, "loss")
    for ga, gb in zip(ra["grads"], rb["grads"]):
        assert_rel(gb, ga, GRAD_RTOL, "grad")


# ------------------------------------------------------------------------------------------------
# tensor-core (tcgen05) adjoint of the neural-ODE shape classes
def _grads_with(case, batch, force):
    import os
    old = os.environ.get("NM_TC_BWD")
    os.environ["NM_TC_BWD"] = force
    try:
        return H.product_run(case, batch)
    finally:
        if old is None:
            os.environ.pop("NM_TC_BWD", None)
        else:
            os.environ["NM_TC_BWD"] = old


def test_tensor_core_adjoint_matches_exact_adjoint_at_scale():
    """c3 runs its weight-gradient GEMMs as single-pass TF32 on round-to-nearest operands (the remainder tiles do
    not fit in shared memory); the rounding noise averages out over the products of each gradient entry.  At the
    batch size where the dispatcher starts to pick it, the tensor-core adjoint must agree with the exact FFMA
    adjoint (itself pinned to the oracle above) far inside the 1e-4 gradient tolerance, and at a small batch it
    must still be within the tolerance against the oracle."""
    case = H.make_case("c3", DEV, seed=5)
    batch_cpu = case.make_batch(4096, "cpu", 11)
    batch = to_dev(batch_cpu)
    exact = _grads_with(case, batch, "0")
    tcore = _grads_with(case, batch, "1")
    for k in exact["traj"]:
        assert torch.equal(exact["traj"][k], tcore["traj"][k])           # same forward kernel
    for i, (a, b) in enumerate(zip(tcore["grads"], exact["grads"])):
        assert_rel(a, b, 2e-5, f"c3 B=4096 tensor-core vs exact adjoint, grad {i}")
    small_cpu = case.make_batch(300, "cpu", 12)
    params_cpu = [p.detach().cpu().clone() for p in case.plan_params]
    got = _grads_with(case, to_dev(small_cpu), "1")
    ref = H.oracle_run("c3", params_cpu, small_cpu)
    for i, (a, b) in enumerate(zip(got["grads"], ref["grads"])):
        assert_rel(a, b, GRAD_RTOL, f"c3 B=300 tensor-core adjoint vs oracle, grad {i}")


def test_tensor_core_adjoint_is_deterministic():
    case = H.make_case("c3", DEV, seed=5)
    batch = to_dev(case.make_batch(1000, "cpu", 13))
    a = _grads_with(case, batch, "1")
    b = _grads_with(case, batch, "1")
    for x, y in zip(a["grads"], b["grads"]):
        assert torch.equal(x, y)


@pytest.mark.parametrize("name,B_scale", [("c2", 4096), ("c1", 4096), ("t_vdp_rk4_tanh_clamp", 4096)])
def test_tensor_core_policy_adjoint(name, B_scale):
    """Policy mode of the tcgen05 kernels (closed-loop MLP policy evaluated once per step, white-box dynamics in
    registers, raw policy output as the checkpoint): the forward kernel is the 3xTF32 one at every batch size (pinned to
    the goldens / oracle by the tests above); its adjoint runs single-pass TF32 weight-gradient GEMMs, so it is
    compared with the exact FFMA adjoint at scale and with the oracle at a small batch."""
    case = H.make_case(name, DEV, seed=5)
    batch = to_dev(case.make_batch(B_scale, "cpu", 11))
    exact = _grads_with(case, batch, "0")
    tcore = _grads_with(case, batch, "1")
    for k in exact["traj"]:
        assert torch.equal(exact["traj"][k], tcore["traj"][k])           # same forward kernel
    for i, (a, b) in enumerate(zip(tcore["grads"], exact["grads"])):
        assert_rel(a, b, 3e-5, f"{name} B={B_scale} tensor-core vs exact adjoint, grad {i}")
    again = _grads_with(case, batch, "1")
    for x, y in zip(tcore["grads"], again["grads"]):
        assert torch.equal(x, y)                                         # fixed-order accumulation
    for B in (129, 300):          # (single-pass TF32 noise ~ 2^-12 / sqrt(#products): the dispatcher never picks it this low)
        small_cpu = case.make_batch(B, "cpu", 12)
        params_cpu = [p.detach().cpu().clone() for p in case.plan_params]
        got = _grads_with(case, to_dev(small_cpu), "1")
        ref = H.oracle_run(name, params_cpu, small_cpu)
        compare(got, ref, f"{name} B={B} tensor-core policy adjoint vs oracle")


def test_device_prefetcher_feeds_the_same_batches():
    from neuromancer_b200.dataset import DevicePrefetcher
    case = H.make_case("c2", DEV)
    host = [{k: (v.pin_memory() if isinstance(v, torch.Tensor) else v) for k, v in case.make_batch(256, "cpu", s).items()} for s in range(4)]
    want = [float(case.problem(to_dev(b))["train_loss"].detach()) for b in host]
    got = [float(case.problem(b)["train_loss"].detach()) for b in DevicePrefetcher(iter(host), DEV)]
    assert got == want


# ------------------------------------------------------------------------------------------------
# BASELINE-size runs through size-independent properties (the oracle cannot run these sizes in seconds)
FULL_SIZE = [("c1", 3333), ("c2", 65536), ("c3", 65536), ("c4p", 32768), ("c5", 4096)]


@pytest.mark.parametrize("name,B", FULL_SIZE)
def test_full_size_shard_additivity_and_determinism(name, B):
    """Trajectories are independent given the weights and every loss term is a mean: the loss and the gradient of a
    batch must equal the average over its two halves (this is also what the data-parallel all-reduce relies on),
    rows of the trajectories must not depend on which launch / tile / wave produced them, and a repeat must be
    bitwise identical.  Runs the persistent kernels over many waves of tiles at the benchmark batch sizes."""
    case = H.make_case(name, DEV, seed=2)
    full = to_dev(case.make_batch(B, "cpu", 21))
    half = B // 2
    lo = {k: (v[:half].contiguous() if isinstance(v, torch.Tensor) else v) for k, v in full.items()}
    hi = {k: (v[half:2 * half].contiguous() if isinstance(v, torch.Tensor) else v) for k, v in full.items()}
    both = {k: (v[:2 * half].contiguous() if isinstance(v, torch.Tensor) else v) for k, v in full.items()}
    r_all = H.product_run(case, both)
    r_again = H.product_run(case, both)
    r_lo, r_hi = H.product_run(case, lo), H.product_run(case, hi)
    for k in r_all["traj"]:
        assert torch.equal(r_all["traj"][k], r_again["traj"][k])
        assert torch.equal(r_all["traj"][k][:half], r_lo["traj"][k]), f"{name}: trajectory rows depend on the launch ({k})"
        assert torch.equal(r_all["traj"][k][half:], r_hi["traj"][k])
        assert torch.isfinite(r_all["traj"][k]).all()
    assert torch.equal(r_all["loss"], r_again["loss"])
    assert_rel(r_all["loss"], 0.5 * (r_lo["loss"] + r_hi["loss"]), 1e-5, f"{name} B={B} loss additivity")
    for i, (g, a, b) in enumerate(zip(r_all["grads"], r_lo["grads"], r_hi["grads"])):
        assert torch.equal(g, r_again["grads"][i])
        assert_rel(g, 0.5 * (a + b), 2e-5, f"{name} B={B} gradient additivity, grad {i}")


def test_white_box_constants_are_refreshed_and_identified():
    """System identification on the fused path (reference ode.py:216-237: c1, c2 are trainable nn.Parameters; examples/ODEs
    Part_1/2 param_estim): (a) an in-place change of a constant is seen by the next rollout (the kernels read the constants
    by value from the plan, refreshed per call); (b) Adam on d loss / d constants, straight from the adjoint kernel, recovers
    the constants that generated the data."""
    case = H.make_case("t_sysid_twotank", DEV)
    c1, c2 = case.plan_params
    batch = to_dev(case.make_batch(512, "cpu", 11))
    with torch.no_grad():
        c1.fill_(0.08), c2.fill_(0.04)
        truth = case.problem(batch)["train_xn"].clone()
        c1.fill_(0.12)
        moved = case.problem(batch)["train_xn"]
        assert not torch.equal(moved, truth)                                   # (a)
        c1.fill_(0.08)
        assert torch.equal(case.problem(batch)["train_xn"], truth)
        batch["X"] = truth[:, :-1, :].contiguous()
        c1.fill_(0.12), c2.fill_(0.07)
    opt = torch.optim.Adam([c1, c2], lr=2e-3)
    first = None
    for _ in range(400):
        opt.zero_grad()
        loss = case.problem(batch)["train_loss"]
        loss.backward()
        opt.step()
        first = float(loss.detach()) if first is None else first
    assert float(loss.detach()) < 1e-3 * first, (first, float(loss.detach()))
    assert abs(float(c1) - 0.08) < 2e-3 and abs(float(c2) - 0.04) < 2e-3, (float(c1), float(c2))     # (b)


def test_frozen_constants_get_no_gradient_entry():
    """A plant whose constants are all frozen has no derived entry in the arena (and keeps the tensor-core adjoint);
    freezing after the first call re-lowers the plan."""
    case = H.make_case("t_sysid_lorenz", DEV)
    batch = to_dev(case.make_batch(64, "cpu", 3))
    out = case.problem(batch)
    out["train_loss"].backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in case.plan_params)
    assert case.plan().c.n_params == 3 and case.plan().c.rhs_mlp.trainable == 1
    frozen = H.make_case("t_auto_lorenz", DEV)
    assert frozen.plan().c.n_params == 0 and frozen.plan().c.rhs_mlp.trainable == 0


@pytest.mark.parametrize("barrier", ["log10", "softexp"])
def test_barrier_loss_is_fused_and_reports_the_reference_keys(barrier):
    """BarrierLoss with a named barrier (reference loss.py:184-257) is lowered: penalty_loss / loss come from the fused
    kernels (no torch-evaluated term), the constraint keys keep the plain weighted penalty, `<name>_barrier` and
    `barrier_loss` materialise on demand and agree with the torch evaluation of the same loss class (callable barrier ->
    not lowered) on the same trajectories -- values and gradients."""
    from neuromancer_b200 import BarrierLoss, Problem
    name = f"t_barrier_{barrier}"
    case = H.make_case(name, DEV, seed=2)
    plan = case.plan()
    assert plan.barrier and not plan.python_terms and plan.c.n_terms == 4
    batch = to_dev(case.make_batch(300, "cpu", 5))
    out = case.problem(batch)
    fused_loss = case.problem.loss
    eager_loss = BarrierLoss(list(fused_loss.objectives), list(fused_loss.constraints), barrier=fused_loss.barriers[barrier],
                             upper_bound=fused_loss.upper_bound, shift=fused_loss.shift, alpha=fused_loss.alpha)
    assert eager_loss.barrier_name is None
    eager = Problem([case.system], eager_loss)
    assert not eager._fusable()
    ref = eager(batch)
    for k in ("train_loss", "train_penalty_loss", "train_objective_loss", "train_barrier_loss"):
        assert_rel(torch.as_tensor(out[k]).detach().cpu(), torch.as_tensor(ref[k]).detach().cpu(), GRAD_RTOL, f"{name} {k}")
    for con in fused_loss.constraints:
        assert_rel(out[f"train_{con.output_keys[0]}"].detach().cpu(), ref[f"train_{con.output_keys[0]}"].detach().cpu(), GRAD_RTOL, con.name)
        assert torch.equal(out[f"train_{con.name}_barrier"], ref[f"train_{con.name}_barrier"])
    params = [p for p in case.plan_params if p.requires_grad]
    g_fused = torch.autograd.grad(out["train_loss"], params)
    g_eager = torch.autograd.grad(ref["train_loss"], params)
    for i, (a, b) in enumerate(zip(g_fused, g_eager)):
        assert_rel(a.cpu(), b.cpu(), GRAD_RTOL, f"{name} grad {i} fused vs torch-evaluated barrier")
