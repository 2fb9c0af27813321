"""
Drop-in for objects built with the REFERENCE's own classes.

The reference has no plugin interface: a ``Problem([System([Node(MLP_bounds), Node(RK4(ode))])], PenaltyLoss(...))`` is
plain Python objects (reference problem.py:159-177, system.py:30-46, 119-136, integrators.py:16-26, blocks.py:131-140,
constraint.py:327-620).  ``from_reference(obj)`` walks such an object graph by DUCK TYPING -- class name, ``linear`` /
``nonlin`` module lists, ``block`` / ``h`` of an integrator, ``ode_equations`` owner class and its constant attributes,
``fx`` / ``fu`` / ``fd`` of an SSM, the ``_func`` / ``_display_name`` / graph edges of a ``Variable`` -- and builds the
equivalent object of this package **around the very same nn.Parameter objects** (an optimiser created on the reference
model keeps training it; ``state_dict`` keys of the adopted model follow the reference's).  The result lowers to the
same POD ``NmPlan`` as a model written against this package (tests/test_adopt_reference.py compares the bytes) and runs
on the fused sm_100a kernels; anything that is not recognised raises ``PlanError`` (no eager fallback).

The stub a reference maintainer would add is in INTEGRATION.md: ``System.forward`` / ``Problem.forward`` try
``neuromancer_b200.adopt.fused(self)`` and fall through to the eager loop when it raises.
"""
from __future__ import annotations

import functools
from typing import Callable, Dict

import torch
import torch.nn as nn

from . import constraint as _c
from .dynamics import integrators as _int
from .dynamics import ode as _ode
from .loss import PenaltyLoss
from .modules import blocks as _blocks
from .plan import PlanError
from .problem import Problem
from .system import Node, System, SystemPreview


def _cls(obj) -> str:
    return type(obj).__name__


def _is_reference(obj) -> bool:
    return type(obj).__module__.split(".")[0] == "neuromancer"


# right-hand sides that are not reference classes (user subclasses of the reference's ODESystem) register a converter
_RHS_CONVERTERS: Dict[str, Callable] = {}


def register_rhs(class_name: str, build: Callable) -> None:
    """``build(reference_rhs) -> neuromancer_b200.dynamics.ode.ODESystem`` for a user-defined ODESystem subclass."""
    _RHS_CONVERTERS[class_name] = build


# ------------------------------------------------------------------------------------------------ blocks
def _inner_linear(layer):
    if isinstance(layer, nn.Linear):
        return layer
    inner = getattr(layer, "linear", None)
    if isinstance(inner, nn.Linear):
        return inner
    raise PlanError(f"linear map {_cls(layer)} is not dense (torch.nn.Linear / slim.Linear): not fused")


def _is_structured(layer) -> bool:
    """a slim map that is not a plain dense layer but can state its matrix (reference slim/linear.py:85-91)"""
    if isinstance(layer, nn.Linear) or isinstance(getattr(layer, "linear", None), nn.Linear):
        return False
    return callable(getattr(layer, "effective_W", None)) and hasattr(layer, "in_features") and hasattr(layer, "out_features")


def _adopt_mlp(ref):
    structured = [_is_structured(l) for l in ref.linear]
    lins = [l if st else _inner_linear(l) for l, st in zip(ref.linear, structured)]
    sizes = [lins[0].in_features] + [l.out_features for l in lins]
    acts = list(ref.nonlin)
    hidden = {type(a) for a in acts[:-1]}
    if len(hidden) > 1 or not isinstance(acts[-1], nn.Identity):
        raise PlanError("MLP with mixed activations / a non-identity output activation is not fused")
    nonlin = type(acts[0]) if len(acts) > 1 else nn.Identity
    bias = lins[0].bias is not None
    kw = dict(insize=sizes[0], outsize=sizes[-1], bias=bias, linear_map=nn.Linear, nonlin=nonlin, hsizes=sizes[1:-1])
    if _cls(ref) == "MLP_bounds":
        mlp = _blocks.MLP_bounds(min=ref.min, max=ref.max, method=getattr(ref, "method_name", None) or _bounds_method(ref), **kw)
    elif _cls(ref) == "MLP":
        mlp = _blocks.MLP(**kw)
    else:
        raise PlanError(f"block {_cls(ref)} is not fused (MLP / MLP_bounds are)")
    # same Parameter objects: the adopted model trains the reference model.  A structured slim map is taken over as the
    # module it is (the lowering materialises its effective_W() on every call)
    for i, st in enumerate(structured):
        if st:
            mlp.linear[i] = lins[i]
    for (w, b), lin, st in zip(mlp.dense_layers(), lins, structured):
        if st:
            continue
        _share(mlp, w, lin.weight)
        if b is not None:
            _share(mlp, b, lin.bias)
    if hasattr(acts[0], "negative_slope") and hasattr(mlp.nonlin[0], "negative_slope"):
        for a, m in zip(acts, mlp.nonlin):
            if hasattr(a, "negative_slope"):
                m.negative_slope = a.negative_slope
    return mlp


def _bounds_method(ref) -> str:
    fn = getattr(ref, "method", None)
    name = getattr(fn, "__name__", str(fn))
    return "sigmoid_scale" if "scaling" in name or "sigmoid" in name else "relu_clamp"


def _share(root: nn.Module, old: nn.Parameter, new: nn.Parameter) -> None:
    """make ``new`` the Parameter object wherever ``old`` is registered under ``root``"""
    for mod in root.modules():
        for k, v in list(mod._parameters.items()):
            if v is old:
                mod._parameters[k] = new


# ------------------------------------------------------------------------------------------------ dynamics
_ODE_BY_NAME = {n: getattr(_ode, n) for n in ("TwoTankParam", "VanDerPolControl", "LorenzControl", "LorenzParam", "CSTR_Param",
                                              "DuffingParam", "BrusselatorParam", "LotkaVolterraParam") if hasattr(_ode, n)}
_INT_BY_NAME = {n: getattr(_int, n) for n in ("Euler", "RK2", "RK4", "Euler_Trap", "RK4_Trap", "Luther", "Runge_Kutta_Fehlberg")
                if hasattr(_int, n)}


def _adopt_rhs(ref):
    name = _cls(ref)
    if name in _RHS_CONVERTERS:
        return _RHS_CONVERTERS[name](ref)
    if name in ("MLP", "MLP_bounds"):
        return _adopt_mlp(ref)
    if name in ("BrusselatorHybrid", "LotkaVolterraHybrid") and hasattr(ref, "block"):       # grey-box: constants around an MLP
        rhs = getattr(_ode, name)(_adopt_mlp(ref.block))
        for k, v in ref._parameters.items():
            if k not in rhs._parameters:
                raise PlanError(f"{name}: constant '{k}' has no counterpart in the fused right-hand side")
            rhs._parameters[k] = v
        return rhs
    if name in _ODE_BY_NAME:
        rhs = _ODE_BY_NAME[name]()
        for k, v in ref._parameters.items():          # constants keep their Parameter objects (c1, c2, mu, ...)
            if k in rhs._parameters:
                rhs._parameters[k] = v
            else:
                raise PlanError(f"{name}: constant '{k}' has no counterpart in the fused right-hand side")
        return rhs
    raise PlanError(f"right-hand side {name} has no fused implementation (register_rhs adds user classes)")


def _adopt_callable(fn):
    name = _cls(fn)
    if name in _INT_BY_NAME and hasattr(fn, "block"):
        return _INT_BY_NAME[name](_adopt_rhs(fn.block), h=fn.h)
    if name == "SSM" and hasattr(fn, "fx") and hasattr(fn, "fu"):
        fd = getattr(fn, "fd", None)
        return _ode.SSM(fn.fx, fn.fu, fn.nx, fn.nu, fd=fd, nd=getattr(fn, "nd", 0) if fd is not None else 0)
    if name in ("MLP", "MLP_bounds"):
        return _adopt_mlp(fn)
    if isinstance(fn, nn.Linear) or isinstance(getattr(fn, "linear", None), nn.Linear):
        return fn                                       # bias-free dense output map y = C x
    raise PlanError(f"node callable {name} is not recognised by the fused rollout (a python function is opaque: express linear "
                    "dynamics / output maps as ode.SSM / nn.Linear modules)")


def _adopt_node(ref):
    return Node(_adopt_callable(ref.callable), list(ref.input_keys), list(ref.output_keys), name=ref.name)


def _adopt_system(ref):
    nodes = [_adopt_node(n) for n in ref.nodes]
    if _cls(ref) == "SystemPreview":
        return SystemPreview(nodes, preview_keys_map=dict(ref.preview_keys_map), preview_length=dict(ref.preview_length),
                             pad_mode=ref.pad_mode, pad_constant=ref.pad_constant, name=ref.name, nstep_key=ref.nstep_key,
                             init_func=None, nsteps=ref.nsteps)
    return System(nodes, name=ref.name, nstep_key=ref.nstep_key, init_func=None, nsteps=ref.nsteps)


# ------------------------------------------------------------------------------------------------ loss terms
_BIN = {"+": "__add__", "-": "__sub__", "∗": "__mul__", "*": "__mul__", "/": "__truediv__", "pow": "__pow__"}


def _adopt_variable(v, memo):
    """reference Variable (networkx DAG of lambdas tagged by display name) -> tagged expression tree of this package"""
    if id(v) in memo:
        return memo[id(v)]
    if getattr(v, "_is_input", False):
        out = _c.variable(v._key)
    elif v._func is None:
        out = v._value                                   # a constant tensor / Parameter node
    elif isinstance(v._func, functools.partial) and not list(v._g.in_edges(v)):
        out = v._func.args[0]                            # python number wrapped as a constant node
    else:
        ins = [_adopt_variable(src, memo) for src, _ in v._g.in_edges(v)]
        tag = v._display_name
        if tag in _BIN and len(ins) == 2:
            a, b = ins
            if isinstance(a, _c.Variable):
                out = getattr(a, _BIN[tag])(b)
            elif isinstance(b, _c.Variable):
                out = getattr(b, "__r" + _BIN[tag][2:])(a)
            else:
                raise PlanError("constant-only expression in a loss term")
        elif tag == "neg" and len(ins) == 1:
            out = -ins[0]
        elif tag == "abs" and len(ins) == 1:
            out = abs(ins[0])
        elif tag == "slice" and len(ins) == 1:
            cells = v._func.__closure__ or ()
            if len(cells) != 1:
                raise PlanError("unrecognised slice node in a loss term")
            out = ins[0][cells[0].cell_contents]
        else:
            # torch function applied to Variables (reference constraint.py:593-604): the closure holds the function, the
            # non-variable arguments (placed first, as the reference does) and the keyword arguments
            names = getattr(getattr(v._func, "__code__", None), "co_freevars", ())
            cells = dict(zip(names, (c.cell_contents for c in (v._func.__closure__ or ()))))
            if {"func", "kwargs", "non_variable_args"} <= set(cells) and any(isinstance(i, _c.Variable) for i in ins):
                out = cells["func"](*cells["non_variable_args"], *ins, **cells["kwargs"])
            else:
                raise PlanError(f"loss-term node '{tag}' (a user lambda) cannot be adopted: build the term with "
                                "neuromancer_b200.variable, which evaluates it with torch ops on the fused trajectories")
    memo[id(v)] = out
    return out


def _adopt_constraint(ref, memo):
    cmp_name = _cls(ref.comparator)
    cmp_cls = {"LT": _c.LT, "GT": _c.GT, "Eq": _c.Eq}.get(cmp_name)
    if cmp_cls is None:
        raise PlanError(f"comparator {cmp_name} is not fused")
    left, right = _adopt_variable(ref.left, memo), _adopt_variable(ref.right, memo)
    con = _c.Constraint(left, right, cmp_cls(norm=ref.comparator.norm), weight=ref.weight, name=ref.name)
    con.update_name(ref.name)
    return con


def _adopt_loss(ref):
    if _cls(ref) != "PenaltyLoss":
        raise PlanError(f"loss {_cls(ref)} is not fused (PenaltyLoss is)")
    memo: dict = {}
    objs = [_adopt_constraint(o, memo) for o in ref.objectives]
    cons = [_adopt_constraint(c, memo) for c in ref.constraints]
    return PenaltyLoss(objs, cons)


# ------------------------------------------------------------------------------------------------ entry points
def from_reference(obj):
    """Equivalent object of this package for a reference-built ``Problem`` / ``System`` / ``SystemPreview`` / ``Node`` /
    ``PenaltyLoss`` / block / integrator, sharing its parameters.  Objects of this package pass through unchanged."""
    if not _is_reference(obj):
        return obj
    name = _cls(obj)
    if name == "Problem":
        nodes = [from_reference(n) for n in obj.nodes]
        return Problem(nodes, _adopt_loss(obj.loss), grad_inference=getattr(obj, "grad_inference", False))
    if name in ("System", "SystemPreview"):
        return _adopt_system(obj)
    if name == "Node":
        return _adopt_node(obj)
    if name == "PenaltyLoss":
        return _adopt_loss(obj)
    return _adopt_callable(obj)


def fused(obj):
    """Cached ``from_reference(obj)`` -- what the INTEGRATION.md stub calls from ``System.forward`` / ``Problem.forward``."""
    cached = obj.__dict__.get("_nm_fused")
    if cached is None:
        cached = from_reference(obj)
        object.__setattr__(obj, "_nm_fused", cached) if not isinstance(obj, nn.Module) else obj.__dict__.__setitem__("_nm_fused", cached)
    return cached
