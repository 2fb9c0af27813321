"""
Plan compiler: lowers a ``System`` (and optionally the ``PenaltyLoss`` scoring it) to the POD
``NmPlan`` descriptor of ``include/nm_rollout.h``.

The reference executes ``System.forward`` as a Python double loop over steps and nodes
(system.py:272-276) calling opaque callables.  Here the node list is *recognised* once:

    [ y = C x ]?   dense bias-free linear map of the state            -> NmPlan.has_output / C
    [ policy ]?    blocks.MLP / blocks.MLP_bounds                      -> NmPlan.policy, segs
    dynamics       integrators.{Euler,RK2,RK4}(ODESystem | MLP)       -> NM_DYN_ODE
                   or ode.SSM with dense bias-free fx, fu, fd         -> NM_DYN_SSM

in exactly that order inside a step (the order fixes which values a node sees).
Anything else raises ``NotImplementedError``: there is no eager fallback for the rollout.

Loss terms (``Constraint`` objects) whose two sides are affine "atoms" of trajectory variables --
``scale * V[:, time-slice, dim-slice] + offset``, optionally combined by one ``+``, ``-`` or ``*``
-- are lowered to ``NmTerm`` and scored inside the kernels.  Other terms stay in Python: they are
evaluated with torch ops on the trajectories the kernel wrote and their gradient w.r.t. the
trajectories is handed to the adjoint kernel (``grad_*_ext`` of ``nm_rollout_backward``).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib as L
from .constraint import Constraint, Variable
from .dynamics import integrators as _int
from .dynamics import ode as _ode
from .modules import blocks as _blocks


class PlanError(NotImplementedError):
    """The node list / term cannot be lowered to the fused sm_100a path."""


# ------------------------------------------------------------------------------- helpers
def _dense_matrix(module, what):
    """``[out, in]`` weight of a bias-free dense linear map, else PlanError."""
    if isinstance(module, _blocks.Linear):
        module = module.linear
    inner = module if isinstance(module, nn.Linear) else getattr(module, "linear", None)
    if not isinstance(inner, nn.Linear):
        raise PlanError(f"{what}: {type(module).__name__} is not a dense linear map; the fused "
                        "path recognises torch.nn.Linear / slim.Linear / blocks.Linear only")
    if inner.bias is not None:
        raise PlanError(f"{what}: linear map must be bias-free (bias=False)")
    return inner.weight


# Tensors whose VALUES are baked into the POD plan (SSM / output-map matrices, white-box ODE constants, bounds, step
# size).  RolloutPlan records (tensor, version counter) for each: an in-place change (optimizer step, load_state_dict,
# .data assignment shows up as a new object) makes the cached plan stale and System.plan_for rebuilds it.
_TRACK: Optional[list] = None


def _track(t):
    if _TRACK is not None and isinstance(t, torch.Tensor):
        _TRACK.append((t, t._version))


def _fill_matrix(dst, weight, rows, cols):
    if weight.requires_grad:
        raise PlanError("a trainable linear map (requires_grad=True) of the SSM / output node is not differentiated by "
                        "the fused path: freeze it (requires_grad_(False)) -- its gradient would silently be zero")
    _track(weight)
    w = weight.detach().to("cpu", torch.float32)
    assert w.shape == (rows, cols), (tuple(w.shape), rows, cols)
    if rows > L.NM_MAX_DIM or cols > L.NM_MAX_DIM:
        raise PlanError(f"linear map {rows}x{cols} exceeds NM_MAX_DIM={L.NM_MAX_DIM}")
    flat = w.reshape(-1).tolist()
    for i, v in enumerate(flat):
        dst[i] = v


def _mlp_descriptor(mlp: _blocks.MLP, what: str) -> Tuple[L.NmMlp, List[torch.nn.Parameter]]:
    """NmMlp (without param_offset) and the parameters in flat order W0, b0, W1, b1, ..."""
    if not isinstance(mlp, _blocks.MLP):
        raise PlanError(f"{what}: expected blocks.MLP / blocks.MLP_bounds, got {type(mlp).__name__}")
    try:
        pairs = mlp.dense_layers()
        act, act_param = mlp.lowered_activation()
        bounds, bmin, bmax = mlp.bound_spec()
    except NotImplementedError as e:
        raise PlanError(f"{what}: {e}") from None
    if len(pairs) > L.NM_MAX_LAYERS:
        raise PlanError(f"{what}: {len(pairs)} layers > NM_MAX_LAYERS={L.NM_MAX_LAYERS}")
    d = L.NmMlp()
    d.n_layers = len(pairs)
    sizes = mlp.layer_sizes()
    for i, s in enumerate(sizes):
        d.sizes[i] = s
    d.act, d.act_param = act, act_param
    has_bias = [b is not None for _, b in pairs]
    if any(has_bias) and not all(has_bias):
        raise PlanError(f"{what}: mixed bias / no-bias layers")
    d.has_bias = int(all(has_bias))
    d.bounds = bounds
    nout = sizes[-1]
    if bounds != L.NM_BOUNDS_NONE:
        if nout > L.NM_MAX_DIM:
            raise PlanError(f"{what}: bounded output width {nout} > NM_MAX_DIM")
        for name, val, dst in (("min", bmin, d.bmin), ("max", bmax, d.bmax)):
            if val is None:
                raise PlanError(f"{what}: bounds need both min and max")
            _track(val)
            t = torch.as_tensor(val, dtype=torch.float32).detach().cpu().reshape(-1)
            if t.numel() == 1:
                t = t.expand(nout)
            if t.numel() != nout:
                raise PlanError(f"{what}: {name} has {t.numel()} entries for {nout} outputs")
            for i in range(nout):
                dst[i] = float(t[i])
    params: List[torch.nn.Parameter] = []
    for w, b in pairs:
        params.append(w)
        if b is not None:
            params.append(b)
    d.trainable = int(any(p.requires_grad for p in params))
    return d, params


# ------------------------------------------------------------------------------- term lowering
@dataclass
class _Atom:
    var: int = L.NM_VAR_CONST
    t_start: int = 0
    t_len: int = 1
    d_start: int = 0
    d_len: int = 1
    scale: float = 1.0
    offset: float = 0.0
    # full (unsliced) shape bookkeeping while lowering
    shape: Optional[Tuple[int, int]] = None   # (T, D) of the sliced view, None for a scalar const


def _const_scalar(node: Variable) -> Optional[float]:
    """Python float if ``node`` is a constant with a single distinct finite value."""
    if node.op == "const" or (node.op not in ("input",) and node._func is None and not node._is_input):
        v = node._value
        if isinstance(v, nn.Parameter) and v.requires_grad:
            return None
        if isinstance(v, torch.Tensor):
            if v.numel() == 0:
                return None
            flat = v.detach().reshape(-1).to(torch.float32).cpu()
            if bool((flat == flat[0]).all()):
                return float(flat[0])
            return None
        if isinstance(v, (int, float)) and not isinstance(v, bool):
            return float(v)
    if node.op == "call" and node.op_args in (torch.zeros_like, torch.ones_like):
        return 0.0 if node.op_args is torch.zeros_like else 1.0
    return None


def _norm_index(idx, length):
    """Normalise one axis index to (start, count) keeping the axis, or None if unsupported."""
    if isinstance(idx, slice):
        start, stop, step = idx.indices(length)
        if step != 1:
            return None
        return start, max(0, stop - start)
    if isinstance(idx, (list, tuple)) and len(idx) == 1 and isinstance(idx[0], int):
        i = idx[0] + length if idx[0] < 0 else idx[0]
        if 0 <= i < length:
            return i, 1
    return None


def _lower_atom(node: Variable, var_ids: Dict[str, int], var_shapes: Dict[str, Tuple[int, int]]) -> Optional[_Atom]:
    """Try to express ``node`` as ``scale * V[slice] + offset``."""
    c = _const_scalar(node)
    if c is not None:
        return _Atom(var=L.NM_VAR_CONST, offset=c)
    if node.op == "input":
        if node.key not in var_ids:
            return None
        T, D = var_shapes[node.key]
        return _Atom(var=var_ids[node.key], t_start=0, t_len=T, d_start=0, d_len=D, shape=(T, D))
    kids = list(node.children_vars)
    if node.op == "getitem" and len(kids) == 1:
        base = _lower_atom(kids[0], var_ids, var_shapes)
        if base is None or base.var == L.NM_VAR_CONST or base.shape is None:
            return None
        idx = node.op_args
        if not isinstance(idx, tuple):
            idx = (idx,)
        if any(i is Ellipsis for i in idx):
            return None
        idx = tuple(idx) + (slice(None),) * (3 - len(idx))
        if len(idx) != 3 or idx[0] != slice(None):
            return None
        T, D = base.shape
        ts, ds = _norm_index(idx[1], T), _norm_index(idx[2], D)
        if ts is None or ds is None or ts[1] == 0 or ds[1] == 0:
            return None
        return _Atom(var=base.var, t_start=base.t_start + ts[0], t_len=ts[1],
                     d_start=base.d_start + ds[0], d_len=ds[1], scale=base.scale,
                     offset=base.offset, shape=(ts[1], ds[1]))
    if node.op == "neg" and len(kids) == 1:
        a = _lower_atom(kids[0], var_ids, var_shapes)
        if a is None:
            return None
        return _Atom(a.var, a.t_start, a.t_len, a.d_start, a.d_len, -a.scale, -a.offset, a.shape)
    if node.op in ("add", "sub", "mul", "div") and len(kids) == 2:
        a, b = (_lower_atom(k, var_ids, var_shapes) for k in kids)
        if a is None or b is None:
            return None
        a_const, b_const = a.var == L.NM_VAR_CONST, b.var == L.NM_VAR_CONST
        if a_const and b_const:
            fa, fb = a.offset, b.offset
            val = {"add": fa + fb, "sub": fa - fb, "mul": fa * fb}.get(node.op)
            if node.op == "div":
                val = fa / fb if fb != 0 else None
            return None if val is None else _Atom(var=L.NM_VAR_CONST, offset=float(val))
        if not (a_const or b_const):
            return None           # two variables: handled one level up as an NmSide
        v, k, k_first = (b, a.offset, True) if a_const else (a, b.offset, False)
        # only fold forms that stay a single fused multiply-add of the loaded value
        if node.op == "add" and v.offset == 0.0:
            return _Atom(v.var, v.t_start, v.t_len, v.d_start, v.d_len, v.scale, k, v.shape)
        if node.op == "sub" and v.offset == 0.0:
            if k_first:   # k - v
                return _Atom(v.var, v.t_start, v.t_len, v.d_start, v.d_len, -v.scale, k, v.shape)
            return _Atom(v.var, v.t_start, v.t_len, v.d_start, v.d_len, v.scale, -k, v.shape)
        if node.op == "mul" and v.offset == 0.0 and v.scale == 1.0:
            return _Atom(v.var, v.t_start, v.t_len, v.d_start, v.d_len, k, 0.0, v.shape)
        return None
    return None


def _lower_side(node: Variable, var_ids, var_shapes) -> Optional[Tuple[int, _Atom, _Atom]]:
    a = _lower_atom(node, var_ids, var_shapes)
    if a is not None:
        return L.NM_OP_NONE, a, _Atom()
    kids = list(node.children_vars)
    ops = {"add": L.NM_OP_ADD, "sub": L.NM_OP_SUB, "mul": L.NM_OP_MUL}
    if node.op in ops and len(kids) == 2:
        a, b = (_lower_atom(k, var_ids, var_shapes) for k in kids)
        if a is not None and b is not None:
            return ops[node.op], a, b
    return None


def _broadcast(shapes):
    T = D = 1
    for s in shapes:
        if s is None:
            continue
        if s[0] != 1:
            if T != 1 and T != s[0]:
                return None
            T = s[0]
        if s[1] != 1:
            if D != 1 and D != s[1]:
                return None
            D = s[1]
    return T, D


def lower_term(term, is_objective: bool, var_ids, var_shapes) -> Optional[L.NmTerm]:
    """``NmTerm`` for a ``Constraint`` or None when it has to stay in Python."""
    if not isinstance(term, Constraint):
        return None
    cmp_id = {"eq": L.NM_CMP_EQ, "lt": L.NM_CMP_LT, "gt": L.NM_CMP_GT}.get(str(term.comparator))
    norm = getattr(term.comparator, "norm", None)
    if cmp_id is None or norm not in (1, 2):
        return None
    try:
        weight = float(term.weight)
    except (TypeError, ValueError):
        return None
    if isinstance(term.weight, torch.Tensor) and term.weight.requires_grad:
        return None
    lhs = _lower_side(term.left, var_ids, var_shapes)
    rhs = _lower_side(term.right, var_ids, var_shapes)
    if lhs is None or rhs is None:
        return None
    atoms = [lhs[1], lhs[2], rhs[1], rhs[2]]
    used = [a for a, live in zip(atoms, (True, lhs[0] != L.NM_OP_NONE, True, rhs[0] != L.NM_OP_NONE)) if live]
    shape = _broadcast([a.shape for a in used])
    if shape is None:
        return None
    if all(a.var == L.NM_VAR_CONST for a in used):
        return None
    out = L.NmTerm()
    out.cmp, out.norm, out.weight, out.is_objective = cmp_id, int(norm), weight, int(is_objective)
    out.t_len, out.d_len = shape
    for side, (op, a, b) in ((out.lhs, lhs), (out.rhs, rhs)):
        side.op = op
        for dst, src in ((side.a, a), (side.b, b)):
            dst.var, dst.t_start, dst.t_len = src.var, src.t_start, src.t_len
            dst.d_start, dst.d_len, dst.scale, dst.offset = src.d_start, src.d_len, src.scale, src.offset
    return out


# ------------------------------------------------------------------------------- system lowering
@dataclass
class SystemLayout:
    """Result of recognising a node list (shape-independent part of the plan)."""
    state_key: str = ""
    action_key: Optional[str] = None
    output_key: Optional[str] = None
    exo_keys: List[str] = field(default_factory=list)
    dist_key: Optional[str] = None
    act_exo_key: Optional[str] = None      # open loop: the action is read from this data key (no policy)
    policy: Optional[_blocks.MLP] = None
    rhs_mlp: Optional[_blocks.MLP] = None
    policy_inputs: List[str] = field(default_factory=list)
    policy_preview: Dict[str, int] = field(default_factory=dict)   # input key -> preview length L
    pad_mode: int = 0
    pad_value: float = 0.0
    dynamics: object = None
    output_map: object = None
    nx: int = 0
    nu: int = 0
    ny: int = 0


def _classify(node):
    fn = node.callable
    if isinstance(fn, _int.Integrator) or isinstance(fn, _ode.SSM):
        return "dynamics"
    if isinstance(fn, _blocks.MLP):
        return "policy"
    try:
        _dense_matrix(fn, "output map")
        return "output"
    except PlanError:
        pass
    raise PlanError(
        f"node '{node.name}': callable {type(fn).__name__} is not recognised by the fused rollout "
        "(expected blocks.MLP/MLP_bounds, integrators.Euler/RK2/RK4, ode.SSM, or a bias-free dense "
        "linear output map). neuromancer_b200 has no eager fallback for System.forward.")


_PAD_MODES = {"constant": L.NM_PAD_CONSTANT, "replicate": L.NM_PAD_REPLICATE, "circular": L.NM_PAD_CIRCULAR,
              "reflect": L.NM_PAD_REFLECT}


def recognise(nodes, preview=None) -> SystemLayout:
    kinds = [_classify(n) for n in nodes]
    if kinds.count("dynamics") != 1:
        raise PlanError(f"a fused System needs exactly one dynamics node, found {kinds.count('dynamics')}")
    if kinds.count("policy") > 1 or kinds.count("output") > 1:
        raise PlanError("at most one policy node and one output-map node are supported")
    expected = [k for k in ("output", "policy", "dynamics") if k in kinds]
    if kinds != expected:
        raise PlanError(f"node order {kinds} is not [output-map?, policy?, dynamics]")
    lay = SystemLayout()
    dyn = nodes[kinds.index("dynamics")]
    if len(dyn.output_keys) != 1 or dyn.output_keys[0] != dyn.input_keys[0]:
        raise PlanError("the dynamics node must map [state, action?, disturbance?] -> [state]")
    lay.state_key, lay.dynamics = dyn.output_keys[0], dyn.callable
    fn = dyn.callable
    lay.nx = int(fn.out_features) * (2 if getattr(fn, "nm_second_order", False) else 1)     # LeapFrog / Yoshida4: X = [x, dx]

    pol = nodes[kinds.index("policy")] if "policy" in kinds else None
    out = nodes[kinds.index("output")] if "output" in kinds else None
    if out is not None:
        if out.input_keys != [lay.state_key] or len(out.output_keys) != 1:
            raise PlanError("the output map must be y = C x of the state key")
        lay.output_key, lay.output_map = out.output_keys[0], out.callable
        lay.ny = int(_dense_matrix(out.callable, "output map").shape[0])
    if pol is not None:
        if len(pol.output_keys) != 1:
            raise PlanError("the policy node must have a single output key")
        lay.action_key, lay.policy, lay.policy_inputs = pol.output_keys[0], pol.callable, list(pol.input_keys)
        lay.nu = int(pol.callable.out_features)
        if lay.action_key in (lay.state_key, lay.output_key):
            raise PlanError("policy output key clashes with the state / output key")

    extra = list(dyn.input_keys[1:])
    if lay.action_key is not None:
        if not extra or extra[0] != lay.action_key:
            raise PlanError("the dynamics node must take the policy's action as its second input")
        extra = extra[1:]
    if isinstance(fn, _ode.SSM):
        if len(extra) > 1:
            raise PlanError("SSM dynamics take at most one disturbance input")
        if extra:
            lay.dist_key = extra[0]
        if lay.action_key is None:
            raise PlanError("SSM dynamics need an action input produced by a policy node")
    else:
        if extra and lay.action_key is None and len(extra) == 1:
            # open-loop rollout: the control input of the integrator comes straight from the data
            lay.act_exo_key = extra[0]
            lay.action_key = None
            extra = []
        if extra:
            raise PlanError(f"integrator dynamics with extra inputs {extra} are not fused")
        if isinstance(fn.block, _blocks.MLP):
            lay.rhs_mlp = fn.block

    if preview is not None:
        keys_map, lengths, mode, constant = preview
        if mode not in _PAD_MODES:
            raise PlanError(f"pad_mode '{mode}' is not one of {sorted(_PAD_MODES)}")
        lay.pad_mode, lay.pad_value = _PAD_MODES[mode], float(constant or 0.0)
        for key, names in keys_map.items():
            for node in nodes:
                if node.name not in names or key not in node.input_keys:
                    continue
                if node is not pol:
                    raise PlanError(f"preview of '{key}' by node '{node.name}': only the policy node may preview")
                if key in (lay.state_key, lay.output_key, lay.action_key):
                    raise PlanError(f"'{key}' is produced by the rollout and cannot be previewed")
                if lengths.get(key) is None:
                    raise PlanError(f"preview_length['{key}'] is not set")
                lay.policy_preview[key] = int(lengths[key])

    def add_exo(key):
        if key not in lay.exo_keys:
            lay.exo_keys.append(key)
    for k in lay.policy_inputs:
        if k not in (lay.state_key, lay.output_key):
            if k == lay.action_key:
                raise PlanError("a policy reading its own previous action is not fused")
            add_exo(k)
    if lay.dist_key is not None:
        add_exo(lay.dist_key)
    if lay.act_exo_key is not None:
        add_exo(lay.act_exo_key)
    return lay


class RolloutPlan:
    """A ``System`` (+ optional ``PenaltyLoss``) lowered for one set of tensor shapes."""

    def __init__(self, layout: SystemLayout, nsteps: int, shapes: Dict[str, Tuple[int, ...]],
                 loss=None):
        self.layout, self.nsteps, self.loss = layout, int(nsteps), loss
        lay = layout
        global _TRACK
        _TRACK = self._tracked = []
        try:
            self._build(lay, shapes, loss)
        finally:
            _TRACK = None
        self._term_weights = self._current_term_weights()

    def _current_term_weights(self):
        extra = (float(self.loss.alpha), float(self.loss.shift), float(self.loss.upper_bound)) if getattr(self, "barrier", 0) else ()
        return tuple(float(t.weight) for t, _ in getattr(self, "fused_terms", [])) + extra

    def values_stale(self) -> bool:
        """True when a tensor or term weight whose value the plan carries has changed since lowering."""
        for t, v in self._tracked:
            if t._version != v:
                return True
        return self._current_term_weights() != self._term_weights

    def _build(self, lay, shapes, loss):
        p = L.NmPlan()
        p.abi_version = L.NM_ABI_VERSION
        p.nx, p.nu, p.ny, p.nsteps = lay.nx, lay.nu, lay.ny, self.nsteps
        if max(lay.nx, lay.nu, lay.ny) > L.NM_MAX_DIM:
            raise PlanError(f"nx/nu/ny larger than NM_MAX_DIM={L.NM_MAX_DIM}")
        if self.nsteps < 1:
            raise PlanError("nsteps must be >= 1 for the fused rollout")
        self.params: List[torch.nn.Parameter] = []
        self.param_slices: List[Tuple[int, int]] = []
        offset = 0

        def adopt(desc, plist):
            nonlocal offset
            desc.param_offset = offset
            for q in plist:
                self.params.append(q)
                self.param_slices.append((offset, q.numel()))
                offset += q.numel()

        # exogenous inputs --------------------------------------------------------------
        if len(lay.exo_keys) > L.NM_MAX_EXO:
            raise PlanError(f"{len(lay.exo_keys)} exogenous inputs > NM_MAX_EXO")
        p.n_exo = len(lay.exo_keys)
        for i, k in enumerate(lay.exo_keys):
            if k not in shapes:
                raise KeyError(k)
            shp = shapes[k]
            if len(shp) != 3:
                raise PlanError(f"input '{k}' must be [batch, time, dim], got {shp}")
            if shp[1] < self.nsteps:     # previewed signals too: the window at step i starts at entry i
                raise PlanError(f"input '{k}' has {shp[1]} time entries < nsteps={self.nsteps}")
            p.exo[i].width, p.exo[i].steps = shp[2], shp[1]

        # policy ------------------------------------------------------------------------
        p.has_policy = int(lay.policy is not None)
        if lay.policy is not None:
            desc, plist = _mlp_descriptor(lay.policy, "policy")
            adopt(desc, plist)
            p.policy = desc
            if len(lay.policy_inputs) > L.NM_MAX_SEGS:
                raise PlanError("too many policy inputs")
            p.n_segs = len(lay.policy_inputs)
            width = 0
            for i, k in enumerate(lay.policy_inputs):
                seg = p.segs[i]
                if k == lay.state_key:
                    seg.kind, seg.index, seg.width = L.NM_SEG_STATE, 0, lay.nx
                elif k == lay.output_key:
                    seg.kind, seg.index, seg.width = L.NM_SEG_OUTPUT, 0, lay.ny
                elif k in lay.policy_preview:
                    j = lay.exo_keys.index(k)
                    plen = lay.policy_preview[k]
                    if plen < 0:
                        raise PlanError("preview length must be >= 0")
                    if lay.pad_mode == L.NM_PAD_REFLECT and plen >= p.exo[j].steps:
                        raise PlanError("reflect padding needs preview_length < signal length")
                    if lay.pad_mode == L.NM_PAD_CIRCULAR and plen > p.exo[j].steps:
                        raise PlanError("circular padding needs preview_length <= signal length")
                    seg.kind, seg.index, seg.width, seg.preview_len = L.NM_SEG_PREVIEW, j, p.exo[j].width * (plen + 1), plen
                else:
                    j = lay.exo_keys.index(k)
                    seg.kind, seg.index, seg.width = L.NM_SEG_EXO, j, p.exo[j].width
                width += seg.width
            if width != desc.sizes[0]:
                raise PlanError(f"policy insize {desc.sizes[0]} != concatenated input width {width}")
            if desc.sizes[desc.n_layers] != lay.nu:
                raise PlanError("policy outsize mismatch")

        # output map --------------------------------------------------------------------
        p.has_output = int(lay.output_map is not None)
        if lay.output_map is not None:
            _fill_matrix(p.C, _dense_matrix(lay.output_map, "output map"), lay.ny, lay.nx)

        # dynamics ----------------------------------------------------------------------
        fn = lay.dynamics
        p.dist_exo = -1
        p.act_exo = -1
        if lay.act_exo_key is not None:
            j = lay.exo_keys.index(lay.act_exo_key)
            p.act_exo = j
            lay.nu = p.nu = p.exo[j].width
        if isinstance(fn, _ode.SSM):
            p.dyn_kind, p.integrator, p.rhs_kind, p.h = L.NM_DYN_SSM, 0, L.NM_RHS_NONE, 1.0
            _fill_matrix(p.A, _dense_matrix(fn.fx, "SSM.fx"), lay.nx, lay.nx)
            _fill_matrix(p.Bm, _dense_matrix(fn.fu, "SSM.fu"), lay.nx, lay.nu)
            if lay.dist_key is not None:
                if fn.fd is None:
                    raise PlanError("SSM was given a disturbance input but has no fd map")
                j = lay.exo_keys.index(lay.dist_key)
                p.dist_exo = j
                _fill_matrix(p.E, _dense_matrix(fn.fd, "SSM.fd"), lay.nx, p.exo[j].width)
        else:
            if fn.nm_integrator_id is None:
                raise PlanError(f"integrator {type(fn).__name__} is not fused (Euler, RK2, RK4 are)")
            p.dyn_kind, p.integrator, p.h = L.NM_DYN_ODE, fn.nm_integrator_id, fn.step_size()
            _track(fn.h)
            rhs = fn.block
            if isinstance(rhs, _blocks.MLP):
                p.rhs_kind = L.NM_RHS_MLP
                desc, plist = _mlp_descriptor(rhs, "MLP right-hand side")
                if desc.bounds != L.NM_BOUNDS_NONE:
                    raise PlanError("a bounded MLP as ODE right-hand side is not fused")
                if getattr(fn, "nm_second_order", False):
                    # LeapFrog / Yoshida4 (reference integrators.py:340-404): X = [x, dx], the block is ddx = f(x[, u])
                    if lay.nx % 2 or desc.sizes[0] != lay.nx // 2 + lay.nu or desc.sizes[desc.n_layers] != lay.nx // 2:
                        raise PlanError(f"{type(fn).__name__}: the MLP right-hand side must map [x (nx/2), u] -> ddx (nx/2)")
                elif desc.sizes[0] != lay.nx + lay.nu or desc.sizes[desc.n_layers] != lay.nx:
                    raise PlanError("MLP right-hand side must map [x, u] -> dx")
                adopt(desc, plist)
                p.rhs_mlp = desc
            elif isinstance(rhs, _ode.HybridODESystem):
                # grey-box right-hand side (reference ode.py:284-331): white-box equations around one unknown term
                # m = block(x), an MLP 2 -> 1 evaluated by the kernels' right-hand-side stage loops
                if getattr(fn, "nm_second_order", False):
                    raise PlanError(f"{type(fn).__name__} integrates ddx = f(x): the hybrid ODE systems are first-order")
                if lay.nx != 2 or lay.nu != 0:
                    raise PlanError(f"{type(rhs).__name__} is an autonomous two-state system, got nx={lay.nx}, nu={lay.nu}")
                if not isinstance(rhs.block, _blocks.MLP):
                    raise PlanError(f"{type(rhs).__name__}: the unknown term must be a blocks.MLP, got {type(rhs.block).__name__}")
                desc, plist = _mlp_descriptor(rhs.block, f"{type(rhs).__name__}.block")
                if desc.bounds != L.NM_BOUNDS_NONE or desc.sizes[0] != 2 or desc.sizes[desc.n_layers] != 1:
                    raise PlanError(f"{type(rhs).__name__}.block must be an unbounded MLP mapping x (2) -> 1")
                p.rhs_kind = rhs.nm_rhs_id
                for i, v in enumerate(rhs.rhs_constants()):
                    p.rhs_const[i] = v
                own = [q for n, q in rhs.named_parameters() if not n.startswith("block.")]
                theta = any(q.requires_grad for q in own)
                if theta:
                    plist = plist + [_ode.DerivedConstants(rhs)]
                else:
                    for q in own:
                        _track(q)
                adopt(desc, plist)
                desc.trainable = int(desc.trainable) | (2 if theta else 0)      # bit 0: MLP weights, bit 1: constants
                p.rhs_mlp = desc
            elif isinstance(rhs, _ode.ODESystem) and rhs.nm_rhs_id is not None:
                if getattr(fn, "nm_second_order", False):
                    raise PlanError(f"{type(fn).__name__} integrates ddx = f(x): the white-box ODE systems are first-order")
                p.rhs_kind = rhs.nm_rhs_id
                consts = rhs.rhs_constants()
                for i, v in enumerate(consts):
                    p.rhs_const[i] = v
                if any(q.requires_grad for q in rhs.parameters()):
                    # system identification (reference ode.py:221-225, 357-361: the constants are trainable nn.Parameters):
                    # the constant vector is a derived entry of the flat arena -- its values are refreshed into rhs_const on
                    # every call (ParamArena.materialise), d loss / d rhs_const is written at rhs_mlp.param_offset by the
                    # exact FFMA adjoint and autograd carries it on to the module's parameters
                    p.rhs_mlp.trainable = 1
                    adopt(p.rhs_mlp, [_ode.DerivedConstants(rhs)])
                else:
                    for q in rhs.parameters():
                        _track(q)
                if int(rhs.nu) != lay.nu:
                    raise PlanError(f"{type(rhs).__name__} expects nu={rhs.nu}, policy gives {lay.nu}")
            else:
                raise PlanError(f"right-hand side {type(rhs).__name__} has no fused implementation")
        p.n_params = offset
        p.pad_mode, p.pad_value = lay.pad_mode, lay.pad_value

        # loss terms --------------------------------------------------------------------
        self.var_ids: Dict[str, int] = {lay.state_key: L.NM_VAR_X}
        self.var_shapes: Dict[str, Tuple[int, int]] = {lay.state_key: (self.nsteps + 1, lay.nx)}
        if lay.action_key is not None:
            self.var_ids[lay.action_key] = L.NM_VAR_U
            self.var_shapes[lay.action_key] = (self.nsteps, lay.nu)
        if lay.output_key is not None:
            self.var_ids[lay.output_key] = L.NM_VAR_Y
            self.var_shapes[lay.output_key] = (self.nsteps, lay.ny)
        self.exo_keys = list(lay.exo_keys)
        if loss is not None:
            # data keys only the loss reads (e.g. the measured trajectory 'X' of a system-ID
            # problem) are passed as additional exogenous inputs for the term evaluation
            for k in loss.input_keys:
                if k in self.var_ids or k in self.exo_keys or k not in shapes:
                    continue
                shp = shapes[k]
                if len(shp) == 3 and len(self.exo_keys) < L.NM_MAX_EXO:
                    j = len(self.exo_keys)
                    self.exo_keys.append(k)
                    p.exo[j].width, p.exo[j].steps = shp[2], shp[1]
            p.n_exo = len(self.exo_keys)
        for i, k in enumerate(self.exo_keys):
            self.var_ids[k] = L.NM_VAR_EXO0 + i
            self.var_shapes[k] = (p.exo[i].steps, p.exo[i].width)
        self.fused_terms: List[Tuple[object, bool]] = []      # (term, is_objective)
        self.python_terms: List[Tuple[object, bool]] = []
        # BarrierLoss (reference loss.py:184-257) with one of its named barrier functions: the constraints are lowered with
        # the barrier kind in NmTerm.norm, its parameters travel in the plan; a callable barrier stays in Python
        barrier = L.NM_BARRIERS.get(getattr(loss, "barrier_name", None), 0) if loss is not None else 0
        if barrier:
            p.barrier_alpha, p.barrier_shift, p.barrier_ub = float(loss.alpha), float(loss.shift), float(loss.upper_bound)
        self.barrier = barrier
        if loss is not None:
            listed = [(t, True) for t in loss.objectives] + [(t, False) for t in loss.constraints]
            for term, is_obj in listed:
                low = None
                if len(self.fused_terms) < L.NM_MAX_TERMS:
                    low = lower_term(term, is_obj, self.var_ids, self.var_shapes)
                    if low is not None and barrier and not is_obj:
                        low.norm |= barrier << 4          # NM_TERM_BARRIER: satisfied entries contribute the barrier value
                if low is None:
                    self.python_terms.append((term, is_obj))
                else:
                    p.terms[len(self.fused_terms)] = low
                    self.fused_terms.append((term, is_obj))
            p.n_terms = len(self.fused_terms)
        self.c = p
        self._signature = None

    # ---------------------------------------------------------------------------------
    @property
    def n_scalars(self):
        return self.c.n_terms + 3

    def signature(self) -> str:
        if self._signature is None:
            self._signature = L.signature(self.c)
        return self._signature

    def describe(self) -> str:
        p = self.c
        pol = [p.policy.sizes[i] for i in range(p.policy.n_layers + 1)] if p.has_policy else None
        rhs = [p.rhs_mlp.sizes[i] for i in range(p.rhs_mlp.n_layers + 1)] if L.rhs_has_mlp(p.rhs_kind) else None
        return (f"RolloutPlan(nx={p.nx}, nu={p.nu}, ny={p.ny}, nsteps={p.nsteps}, policy={pol}, "
                f"dyn={p.dyn_kind}, integ={p.integrator}, rhs={p.rhs_kind}, rhs_mlp={rhs}, "
                f"exo={[(e.width, e.steps) for e in p.exo[:p.n_exo]]}, fused_terms={p.n_terms}, "
                f"python_terms={len(self.python_terms)})")
