"""
Bridge between the Python API and the C ABI (``libnm_rollout.so``).

The two entry points are registered with ``torch.library`` as the custom operators ``nmfused::rollout_fwd``
(``nm_rollout_forward``: one persistent kernel for the whole horizon + a tiny finalise kernel) and
``nmfused::rollout_bwd`` (``nm_rollout_backward``: the fused adjoint / BPTT kernels + gradient finalise), tied together
by ``register_autograd``.  PyTorch is used for
device memory, streams and autograd bookkeeping only; all arithmetic of the path happens in the
sm_100a kernels under ``csrc/``.

No fallback: non-CUDA tensors or a missing extension raise immediately.
"""
from __future__ import annotations

import ctypes as C
import weakref
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib as L
from .plan import PlanError, RolloutPlan
from .dynamics.ode import DerivedConstants
from .slim.linear import DerivedWeight

_WORLD = {"size": 1}   # set by neuromancer_b200.distributed when gradients are all-reduced


def set_world_size(n: int) -> None:
    """Loss means are taken over ``B_local * world_size`` trajectories in the adjoint kernel, so a
    SUM all-reduce of the flat gradient equals the global-mean gradient (DDP-average semantics)."""
    _WORLD["size"] = int(n)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(
            f"neuromancer_b200: '{what}' lives on {t.device}; the fused rollout runs on CUDA (sm_100a) "
            "only and has no CPU fallback. Move the batch and the problem to a CUDA device.")
    if t.dtype != torch.float32:
        raise RuntimeError(f"neuromancer_b200: '{what}' must be float32, got {t.dtype}")


class ParamArena:
    """Flat fp32 buffer the kernels read parameters from / write gradients to.

    Module parameters are re-pointed to be views of the arena (``p.data = arena[off:off+n]``), so
    optimiser updates and ``load_state_dict`` write straight into the buffer the kernel reads and
    no per-step packing is needed.  If something replaced ``p.data`` (``module.to(...)``) the views
    are re-adopted on the next call."""

    def __init__(self, plan: RolloutPlan, device):
        self.n = int(plan.c.n_params)
        self.flat = torch.empty(max(self.n, 1), dtype=torch.float32, device=device)
        self.slices = plan.param_slices
        self.plan_c = plan.c
        self.adopt(plan.params)

    def adopt(self, params):
        for p, (off, n) in zip(params, self.slices):
            if isinstance(p, DerivedWeight):
                continue
            view = self.flat[off:off + n].view(p.shape)
            if p.data.data_ptr() != view.data_ptr() or p.data.device != self.flat.device:
                view.copy_(p.data.to(self.flat.device, torch.float32))
                p.data = view

    def in_sync(self, params) -> bool:
        base = self.flat.data_ptr()
        return all(isinstance(p, DerivedWeight) or p.data.data_ptr() == base + 4 * off for p, (off, _) in zip(params, self.slices))

    def materialise(self, params) -> List[torch.Tensor]:
        """The tensors of this call, in plan order: the module Parameters themselves (views of the arena) and, for every
        structured slim map, its ``effective_W()^T`` evaluated now -- copied into its arena slot for the kernels, returned
        with its autograd history so that the kernels' gradient w.r.t. it flows on to the map's own parameters."""
        out = []
        for p, (off, n) in zip(params, self.slices):
            if isinstance(p, DerivedWeight):
                w = p.materialise().to(self.flat.device, torch.float32)
                with torch.no_grad():
                    self.flat[off:off + n].copy_(w.reshape(-1))
                if isinstance(p, DerivedConstants):        # the kernels read white-box constants from the plan, by value
                    for i, v in enumerate(w.detach().cpu().reshape(-1).tolist()):
                        self.plan_c.rhs_const[i] = v
                out.append(w)
            else:
                out.append(p)
        return out


class _Workspace:
    def __init__(self):
        self.buf = None

    def get(self, nbytes: int, device) -> torch.Tensor:
        if self.buf is None or self.buf.numel() < nbytes or self.buf.device != device:
            self.buf = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device)
        return self.buf


# ------------------------------------------------------------------------------------------------
# PyTorch binding: the two C-ABI entry points as custom operators
#   nmfused::rollout_fwd   nm_rollout_forward  (persistent rollout kernel + loss finalise)
#   nmfused::rollout_bwd   nm_rollout_backward (adjoint / BPTT kernels + gradient finalise)
# registered with torch.library, differentiable through register_autograd, launched on torch's current stream.  A plan
# is not a tensor: the operators take the handle of its RolloutRunner (POD NmPlan + parameter arena + workspace).
_RUNNERS: Dict[int, "weakref.ReferenceType"] = {}


def _runner(handle: int) -> "RolloutRunner":
    ref = _RUNNERS.get(handle)
    r = ref() if ref is not None else None
    if r is None:
        raise RuntimeError(f"nmfused: rollout plan handle {handle} is not alive")
    return r


def _empty(ref: torch.Tensor) -> torch.Tensor:
    return ref.new_empty(0)


@torch.library.custom_op("nmfused::rollout_fwd", mutates_args=(), device_types="cuda")
def rollout_fwd(handle: int, x0: torch.Tensor, exo: Sequence[torch.Tensor], params: Sequence[torch.Tensor],
                keep_checkpoints: bool) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """-> (x [B, T+1, nx], u [B, T, nu] | empty, y [B, T, ny] | empty, scalars [n_terms + 3], checkpoints | empty).
    `params` are the module parameters (views of the runner's flat arena, which is what the kernels read)."""
    runner = _runner(handle)
    plan: RolloutPlan = runner.plan
    p = plan.c
    lib = L.load()
    B, dev = x0.shape[0], x0.device
    x_traj = torch.empty(B, p.nsteps + 1, p.nx, dtype=torch.float32, device=dev)
    u_traj = torch.empty(B, p.nsteps, p.nu, dtype=torch.float32, device=dev) if p.nu else None
    y_traj = torch.empty(B, p.nsteps, p.ny, dtype=torch.float32, device=dev) if p.ny else None
    scalars = torch.empty(plan.n_scalars, dtype=torch.float32, device=dev)
    fwd_b, bwd_b = C.c_size_t(), C.c_size_t()
    L.check(lib.nm_rollout_workspace_bytes(C.byref(p), B, C.byref(fwd_b), C.byref(bwd_b)), "nm_rollout_workspace_bytes")
    ws = runner.workspace.get(max(fwd_b.value, bwd_b.value), dev)
    exo_ptrs = (C.c_void_p * max(len(exo), 1))(*[e.data_ptr() for e in exo])
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    chk = None
    if keep_checkpoints:                              # a backward pass may follow: keep the checkpoints
        cb = C.c_size_t()
        L.check(lib.nm_rollout_checkpoint_bytes(C.byref(p), B, C.byref(cb)), "nm_rollout_checkpoint_bytes")
        if cb.value:
            chk = torch.empty(cb.value // 4, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.nm_rollout_forward(C.byref(p), _ptr(runner.arena.flat), _ptr(x0), exo_ptrs,
                                       _ptr(x_traj), _ptr(u_traj), _ptr(y_traj), _ptr(scalars), _ptr(chk),
                                       _ptr(ws), ws.numel(), B, stream), "nm_rollout_forward")
    runner.launches += lib.nm_last_launch_count()
    return (x_traj, u_traj if u_traj is not None else _empty(x_traj), y_traj if y_traj is not None else _empty(x_traj),
            scalars, chk if chk is not None else _empty(x_traj))


@rollout_fwd.register_fake
def _(handle, x0, exo, params, keep_checkpoints):
    runner = _runner(handle)
    p, B = runner.plan.c, x0.shape[0]
    lib = L.load()
    n_chk = 0
    if keep_checkpoints:
        cb = C.c_size_t()
        L.check(lib.nm_rollout_checkpoint_bytes(C.byref(p), B, C.byref(cb)), "nm_rollout_checkpoint_bytes")
        n_chk = cb.value // 4
    return (x0.new_empty(B, p.nsteps + 1, p.nx), x0.new_empty((B, p.nsteps, p.nu) if p.nu else (0,)),
            x0.new_empty((B, p.nsteps, p.ny) if p.ny else (0,)), x0.new_empty(runner.plan.n_scalars), x0.new_empty(n_chk))


@torch.library.custom_op("nmfused::rollout_bwd", mutates_args=(), device_types="cuda")
def rollout_bwd(handle: int, exo: Sequence[torch.Tensor], x_traj: torch.Tensor, u_traj: torch.Tensor, y_traj: torch.Tensor,
                checkpoints: torch.Tensor, grad_scalars: Optional[torch.Tensor], grad_x: Optional[torch.Tensor],
                grad_u: Optional[torch.Tensor], grad_y: Optional[torch.Tensor], want_x0: bool) -> Tuple[torch.Tensor, torch.Tensor]:
    """-> (flat parameter gradient [n_params], d loss / d x0 [B, nx] | empty).  grad_scalars: upstream gradient of the
    [terms | objective | penalty | loss] scalars; grad_x/u/y: upstream gradients of the trajectories (torch-evaluated terms)."""
    runner = _runner(handle)
    plan: RolloutPlan = runner.plan
    p = plan.c
    lib = L.load()
    dev, B = x_traj.device, x_traj.shape[0]
    world = _WORLD["size"]

    def ext(g):
        # gradients of torch-evaluated terms (python_terms, BarrierLoss, user losses on the trajectories) are LOCAL means;
        # the fused terms are divided by the global trajectory count in the kernel, and the data-parallel all-reduce
        # SUMS -- scale the external part by 1 / world_size so that both follow DDP-average semantics
        if g is None or g.numel() == 0:
            return None
        g = g.contiguous()
        return g / world if world > 1 else g
    gx, gu, gy = ext(grad_x), ext(grad_u) if p.nu else None, ext(grad_y) if p.ny else None
    gs = grad_scalars.contiguous() if grad_scalars is not None else torch.zeros(plan.n_scalars, dtype=torch.float32, device=dev)
    gflat = torch.empty_like(runner.arena.flat)
    fwd_b, bwd_b = C.c_size_t(), C.c_size_t()
    L.check(lib.nm_rollout_workspace_bytes(C.byref(p), B, C.byref(fwd_b), C.byref(bwd_b)), "nm_rollout_workspace_bytes")
    ws = runner.workspace.get(max(fwd_b.value, bwd_b.value), dev)
    exo_ptrs = (C.c_void_p * max(len(exo), 1))(*[e.data_ptr() for e in exo])
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    gx0 = torch.empty(B, p.nx, dtype=torch.float32, device=dev) if want_x0 else None
    with torch.cuda.device(dev):
        L.check(lib.nm_rollout_backward(
            C.byref(p), _ptr(runner.arena.flat), exo_ptrs, _ptr(x_traj), _ptr(u_traj if u_traj.numel() else None),
            _ptr(y_traj if y_traj.numel() else None), _ptr(checkpoints if checkpoints.numel() else None),
            _ptr(gs), _ptr(gx), _ptr(gu), _ptr(gy), _ptr(gflat), _ptr(gx0), _ptr(ws), ws.numel(),
            B, B * world, stream), "nm_rollout_backward")
    runner.launches += lib.nm_last_launch_count()
    runner.last_grad_flat = gflat
    return gflat, gx0 if gx0 is not None else _empty(gflat)


@rollout_bwd.register_fake
def _(handle, exo, x_traj, u_traj, y_traj, checkpoints, grad_scalars, grad_x, grad_u, grad_y, want_x0):
    runner = _runner(handle)
    return torch.empty_like(runner.arena.flat), x_traj.new_empty((x_traj.shape[0], runner.plan.c.nx) if want_x0 else (0,))


def _fwd_setup_context(ctx, inputs, output):
    handle, x0, exo, params, _keep = inputs
    x_traj, u_traj, y_traj, _scalars, chk = output
    ctx.handle, ctx.n_exo, ctx.n_params = handle, len(exo), len(params)
    ctx.save_for_backward(x_traj, u_traj, y_traj, chk, *exo)
    ctx.mark_non_differentiable(chk)


def _fwd_backward(ctx, gx, gu, gy, gs, _gchk):
    runner = _runner(ctx.handle)
    plan = runner.plan
    saved = ctx.saved_tensors
    x_traj, u_traj, y_traj, chk = saved[:4]
    exo = list(saved[4:])
    want_x0 = bool(ctx.needs_input_grad[1])
    gflat, gx0 = torch.ops.nmfused.rollout_bwd(ctx.handle, exo, x_traj, u_traj, y_traj, chk, gs, gx,
                                               gu if u_traj.numel() else None, gy if y_traj.numel() else None, want_x0)
    grads = [gflat[off:off + n].view(q.shape) if q.requires_grad else None
             for q, (off, n) in zip(plan.params, plan.param_slices)]
    return None, (gx0 if want_x0 else None), [None] * ctx.n_exo, grads, None


rollout_fwd.register_autograd(_fwd_backward, setup_context=_fwd_setup_context)


def _checked_exo(plan: RolloutPlan, data, x_in) -> List[torch.Tensor]:
    """Exogenous / loss-only inputs in plan order.  The kernels index them as ``[B, steps, width]`` with the batch of the
    initial condition: a tensor on another device, with another batch size (incl. a broadcastable batch of 1) or with
    another trailing shape than the plan was lowered for would be read out of bounds -- rejected here instead."""
    out = []
    B = x_in.shape[0]
    for i, k in enumerate(plan.exo_keys):
        t = data[k]
        _require_cuda(t, k)
        if t.device != x_in.device:
            raise PlanError(f"'{k}' lives on {t.device}, the initial condition on {x_in.device}")
        want = (B, int(plan.c.exo[i].steps), int(plan.c.exo[i].width))
        if t.dim() == 3 and t.shape[0] == 1 and B > 1 and tuple(t.shape[1:]) == want[1:]:
            t = t.expand(B, -1, -1)                      # the reference would broadcast a batch of one
        if tuple(t.shape) != want:
            raise PlanError(f"'{k}' must be {list(want)} ([batch, steps, width] of this plan), got {list(t.shape)}")
        out.append(t.contiguous())
    return out


class RolloutRunner:
    """Per-plan runtime state: parameter arena, workspace, launch counter."""

    def __init__(self, plan: RolloutPlan, device):
        self.plan = plan
        self.arena = ParamArena(plan, device)
        self.workspace = _Workspace()
        self.launches = 0
        self.last_grad_flat = None
        self.handle = id(self)
        _RUNNERS[self.handle] = weakref.ref(self, lambda _r, h=self.handle: _RUNNERS.pop(h, None))
        lib = L.load()
        if not lib.nm_has_kernel(C.byref(plan.c)):
            from . import build
            build.ensure_specialisation(plan)     # compiles + nm_load_plugin; raises on failure
            if not lib.nm_has_kernel(C.byref(plan.c)):
                raise L.NmError(f"no sm_100a specialisation for {plan.signature()}")

    def __call__(self, data: Dict[str, torch.Tensor]):
        plan, lay = self.plan, self.plan.layout
        x_in = data[lay.state_key]
        _require_cuda(x_in, lay.state_key)
        if x_in.dim() != 3 or x_in.shape[2] != lay.nx:
            raise PlanError(f"'{lay.state_key}' must be [batch, 1, {lay.nx}], got {tuple(x_in.shape)}")
        if x_in.shape[1] != 1:
            raise PlanError(
                f"'{lay.state_key}' carries {x_in.shape[1]} time entries; the fused rollout takes the "
                "initial condition only ([batch, 1, nx])")
        for k in (lay.action_key, lay.output_key):
            if k is not None and k in data:
                raise PlanError(f"key '{k}' is produced by the rollout but already present in the data")
        x0 = x_in[:, 0, :].contiguous()
        exo = _checked_exo(plan, data, x_in)
        if not self.arena.in_sync(plan.params):
            self.arena.adopt(plan.params)
        call_params = self.arena.materialise(plan.params)
        keep = torch.is_grad_enabled() and (x0.requires_grad or any(q.requires_grad for q in call_params))
        x, u, y, scalars, _chk = torch.ops.nmfused.rollout_fwd(self.handle, x0, exo, call_params, keep)
        return x, (u if lay.action_key is not None else None), (y if lay.output_key is not None else None), scalars


def _runner_for(plan: RolloutPlan, device) -> RolloutRunner:
    r = getattr(plan, "_runner", None)
    if r is None or r.arena.flat.device != device:
        r = RolloutRunner(plan, device)
        plan._runner = r
    return r


def run_rollout(plan: RolloutPlan, data):
    """Execute ``plan`` on ``data``; returns ``(output dict, scalars tensor)``."""
    lay = plan.layout
    runner = _runner_for(plan, data[lay.state_key].device)
    x, u, y, scalars = runner(data)
    out = dict(data)
    out[lay.state_key] = x
    if u is not None:
        out[lay.action_key] = u
    if y is not None:
        out[lay.output_key] = y
    return out, scalars


def constraint_matrices_fused(plan: RolloutPlan, x_traj, u_traj, y_traj, exo) -> Dict[str, torch.Tensor]:
    """``C_values`` / ``C_violations`` and their equality / inequality column partitions (reference loss.py:86-106) of the
    plan's fused constraints, written by ``nm_constraint_matrices`` (C ABI) in one pass over the trajectories."""
    lib, p = L.load(), plan.c
    B, dev = x_traj.shape[0], x_traj.device
    widths = (C.c_int64 * 3)()
    L.check(lib.nm_constraint_matrices(C.byref(p), None, None, None, None, None, None, None, None, None, None, widths, 0, None),
            "nm_constraint_matrices")
    W, Weq, Win = widths[0], widths[1], widths[2]
    mats = {k: torch.empty(B, w, dtype=torch.float32, device=dev)
            for k, w in (("C_values", W), ("C_violations", W), ("C_eq_values", Weq), ("C_eq_violations", Weq),
                         ("C_ineq_values", Win), ("C_ineq_violations", Win))}
    exo = [e.contiguous() for e in exo]
    exo_ptrs = (C.c_void_p * max(len(exo), 1))(*[e.data_ptr() for e in exo])
    args = [_ptr(t.contiguous()) if t is not None else None for t in (x_traj, u_traj, y_traj)]
    with torch.cuda.device(dev):
        L.check(lib.nm_constraint_matrices(C.byref(p), exo_ptrs, *args, _ptr(mats["C_values"]), _ptr(mats["C_violations"]),
                                           _ptr(mats["C_eq_values"]), _ptr(mats["C_eq_violations"]), _ptr(mats["C_ineq_values"]),
                                           _ptr(mats["C_ineq_violations"]), widths, B,
                                           C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "nm_constraint_matrices")
    return mats


def run_problem(system, loss, data):
    """Fused ``Problem.forward``: rollout + PenaltyLoss scoring (reference problem.py:197-202)."""
    from .problem import LazyOutputs
    batch = system.init(dict(data))
    plan = system.plan_for(batch, loss=loss)
    barrier = getattr(plan, "barrier", 0)
    if barrier and any(not is_obj for _, is_obj in plan.python_terms):
        return None                     # a BarrierLoss constraint that is not lowered: the caller evaluates the loss with torch ops
    traj, scalars = run_rollout(plan, batch)
    out = LazyOutputs(traj)
    n = plan.c.n_terms
    obj_total, pen_total = scalars[n], scalars[n + 1]

    def plain_penalty(term):             # BarrierLoss reports the constraint's ordinary weighted penalty under its own key
        def thunk():
            return {term.output_keys[0]: term.evaluate(traj)[0]}
        return thunk
    for i, (term, is_obj) in enumerate(plan.fused_terms):
        if barrier and not is_obj:       # scalars[i] is weight * mean(penalty | barrier): it enters penalty_loss, not this key
            out.register(term.output_keys[:1], plain_penalty(term))
        else:
            out[term.output_keys[0]] = scalars[i]
    # terms that could not be lowered: evaluated with torch ops on the kernel's trajectories; their
    # gradient reaches the adjoint kernel through grad_x_ext / grad_u_ext / grad_y_ext
    per_constraint = {}
    for term, is_obj in plan.python_terms:
        res = term(traj)
        if isinstance(res, torch.Tensor):
            res = {term.output_keys[0]: res}
        out.update(res)
        val = res[term.output_keys[0]]
        if is_obj:
            obj_total = obj_total + val
        else:
            pen_total = pen_total + val
            per_constraint[id(term)] = (res[term.output_keys[1]], res[term.output_keys[2]])
    out["objective_loss"] = obj_total
    out["penalty_loss"] = pen_total
    out["loss"] = scalars[n + 2] if not plan.python_terms else obj_total + pen_total

    # value / violation tensors and the C_* matrices: on demand only
    fused_constraints = [t for t, is_obj in plan.fused_terms if isinstance(t, type(None)) is False]

    def materialise_term(term):
        def thunk():
            _, value, violation = term.evaluate(traj)
            return {term.output_keys[1]: value, term.output_keys[2]: violation}
        return thunk
    for term, _ in plan.fused_terms:
        out.register(term.output_keys[1:], materialise_term(term))

    if barrier:
        # `<name>_barrier` tensors and `barrier_loss` of the reference's BarrierLoss.calculate_constraints: on demand, torch ops
        def barrier_outputs():
            res, b_loss = {}, 0.0
            for con in loss.constraints:
                _, value, _ = con.evaluate(traj)
                cb = loss.barrier(value)
                cb[cb != cb] = 0.0
                cb[cb == float("Inf")] = 0.0
                cb = torch.clamp(cb, min=0.0, max=loss.upper_bound)
                res[f"{con.name}_barrier"] = cb
                mask = value >= 0
                if (~mask).any():
                    b_loss = b_loss + con.weight * torch.mean(~mask * cb)
            res["barrier_loss"] = b_loss
            return res
        out.register([f"{con.name}_barrier" for con in loss.constraints] + ["barrier_loss"], barrier_outputs)

    if len(loss.constraints):
        out_ref = weakref.ref(out)        # no reference cycle out -> thunk -> out: a step's outputs (and the batch tensors they
                                          # carry) are released by refcount, not by the cyclic collector (10 GB per step on c4 / c5)
        all_fused = not any(not is_obj for _, is_obj in plan.python_terms)
        lay = plan.layout

        def c_matrices():
            o = out_ref()
            x_t, u_t, y_t = traj[lay.state_key], (traj[lay.action_key] if lay.action_key else None), (traj[lay.output_key] if lay.output_key else None)
            tracked = torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in (x_t, u_t, y_t))
            if all_fused and not tracked:
                # one pass of nm_constraint_matrices_kernel over the trajectories (values only: under no_grad, or when nothing
                # upstream is differentiated -- with autograd history wanted the torch expressions below are used)
                return constraint_matrices_fused(plan, x_t, u_t, y_t, _checked_exo(plan, batch, batch[lay.state_key]))
            per = []
            for con in loss.constraints:
                if id(con) in per_constraint:
                    per.append(per_constraint[id(con)])
                else:
                    per.append((o[con.output_keys[1]], o[con.output_keys[2]]))
            return loss.constraint_matrices(loss.constraints, per)
        out.register(["C_violations", "C_values", "C_eq_violations", "C_ineq_violations",
                      "C_eq_values", "C_ineq_values"], c_matrices)
    return out


class RawTrainStep:
    """Allocation-free train step straight on the C ABI: ``forward()`` + ``backward()`` on buffers that
    stay resident in HBM.  The gradient buffer carries the loss scalars in its tail
    (``[n_params | n_terms+3]``) so that data-parallel training needs a single all-reduce.
    Used by ``bench.py`` for the kernel-path throughput; the public API path is ``Problem.forward``."""

    def __init__(self, plan: RolloutPlan, batch: Dict[str, torch.Tensor], world_size: int = 1, peer_sync: bool = False):
        lay, p = plan.layout, plan.c
        self.plan, self.lib = plan, L.load()
        x_in = batch[lay.state_key]
        _require_cuda(x_in, lay.state_key)
        dev = x_in.device
        self.dev, self.B, self.world_size = dev, x_in.shape[0], world_size
        self.runner = _runner_for(plan, dev)
        if not self.runner.arena.in_sync(plan.params):
            self.runner.arena.adopt(plan.params)
        self.x0 = x_in[:, 0, :].contiguous()
        self.exo = _checked_exo(plan, batch, x_in)
        B = self.B
        self.x_traj = torch.empty(B, p.nsteps + 1, p.nx, dtype=torch.float32, device=dev)
        self.u_traj = torch.empty(B, p.nsteps, p.nu, dtype=torch.float32, device=dev) if p.nu else None
        self.y_traj = torch.empty(B, p.nsteps, p.ny, dtype=torch.float32, device=dev) if p.ny else None
        n = int(p.n_params)
        self.gbuf = torch.zeros(n + plan.n_scalars, dtype=torch.float32, device=dev)
        self.grad = self.gbuf[:n]
        self.scalars = self.gbuf[n:]
        # data-parallel exchange inside a kernel over NVLink peer memory (distributed.PeerAllReduce): the kernels of a step
        # write scalars and gradients straight into that step's half of the symmetric buffer, `all_reduce()` sums over
        # ranks into gbuf -- no NCCL call, no staging copy
        self.sync = None
        if peer_sync and world_size > 1:
            from .distributed import PeerAllReduce
            self.sync = PeerAllReduce(n + plan.n_scalars, dev)
            self._n = n
        fwd_b, bwd_b = C.c_size_t(), C.c_size_t()
        L.check(self.lib.nm_rollout_workspace_bytes(C.byref(p), B, C.byref(fwd_b), C.byref(bwd_b)),
                "nm_rollout_workspace_bytes")
        self.ws = torch.empty(max(fwd_b.value, bwd_b.value, 256), dtype=torch.uint8, device=dev)
        cb = C.c_size_t()
        L.check(self.lib.nm_rollout_checkpoint_bytes(C.byref(p), B, C.byref(cb)), "nm_rollout_checkpoint_bytes")
        self.chk = torch.empty(cb.value // 4, dtype=torch.float32, device=dev) if cb.value else None
        self.exo_ptrs = (C.c_void_p * max(len(self.exo), 1))(*[e.data_ptr() for e in self.exo])
        self.launches = 0

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)

    def all_reduce(self):
        """SUM over ranks of [gradient | scalars] of the step just computed -> ``gbuf`` (``grad`` / ``scalars`` views)."""
        if self.sync is not None:
            self.sync.all_reduce(self.gbuf)
            self.grad, self.scalars = self.gbuf[:self._n], self.gbuf[self._n:]
            self.launches += 1
        elif self.world_size > 1:
            import torch.distributed as dist
            dist.all_reduce(self.gbuf)

    def forward(self):
        p = self.plan.c
        if self.sync is not None:                  # this step's half of the symmetric buffer receives scalars and gradients
            half = self.sync.half()
            self.grad, self.scalars = half[:self._n], half[self._n:]
        L.check(self.lib.nm_rollout_forward(
            C.byref(p), _ptr(self.runner.arena.flat), _ptr(self.x0), self.exo_ptrs, _ptr(self.x_traj),
            _ptr(self.u_traj), _ptr(self.y_traj), C.c_void_p(self.scalars.data_ptr()), _ptr(self.chk), _ptr(self.ws),
            self.ws.numel(), self.B, self._stream()), "nm_rollout_forward")
        self.launches += self.lib.nm_last_launch_count()

    def backward(self):
        p = self.plan.c
        L.check(self.lib.nm_rollout_backward(
            C.byref(p), _ptr(self.runner.arena.flat), self.exo_ptrs, _ptr(self.x_traj), _ptr(self.u_traj),
            _ptr(self.y_traj), _ptr(self.chk), None, None, None, None, C.c_void_p(self.grad.data_ptr()), None,
            _ptr(self.ws), self.ws.numel(), self.B, self.B * self.world_size, self._stream()),
            "nm_rollout_backward")
        self.launches += self.lib.nm_last_launch_count()
