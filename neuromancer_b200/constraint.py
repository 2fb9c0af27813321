"""
Symbolic variables, comparators and constraints.

API mirror of the reference's ``constraint.py`` (``Variable`` :327-620, ``variable()`` :622-695,
``LT/GT/Eq`` :65-173, ``Constraint`` :231-324, ``Objective`` :176-228, ``Loss`` :24-62) built on a
different representation: every ``Variable`` is a node of an explicit *tagged expression tree*
(``op`` in {"input", "const", "add", "sub", "mul", "div", "neg", "getitem", "pow", "call", ...})
instead of the reference's networkx graph of opaque lambdas.  The tags are what lets
``neuromancer_b200.plan`` lower a ``Constraint`` to an ``NmTerm`` descriptor evaluated inside the
fused sm_100a kernels; anything it cannot lower is still evaluated eagerly (``forward``) on the
trajectories the kernels produced, and its gradient is fed back into the adjoint kernel.

Comparator semantics (what the kernels reproduce):
    LT: value = l - r,  violation = relu(value) [**2 if norm 2], loss = mean(violation)
    GT: value = r - l,  same
    Eq: value = l - r,  norm 1 -> |value| / l1_loss,  norm 2 -> value**2 / mse_loss
"""
from __future__ import annotations

import operator
from typing import Callable, Dict, Iterable, List

import torch
import torch.nn as nn
import torch.nn.functional as F


def _same_device(left, right):
    """Constants built on the host follow the trajectory tensor's device (the reference does this
    with its ``handle_device_placement`` decorator, utils.py:6-22)."""
    if left.device != right.device:
        if left.device.type == "cpu":
            left = left.to(right.device)
        elif right.device.type == "cpu":
            right = right.to(left.device)
    return left, right


class _Comparator(nn.Module):
    tag = "?"

    def __init__(self, norm=1):
        super().__init__()
        self.norm = norm

    def __str__(self):
        return self.tag


class LT(_Comparator):
    """``left <= right``."""
    tag = "lt"

    def forward(self, left, right):
        left, right = _same_device(left, right)
        value = left - right
        violation = F.relu(value)
        if self.norm == 2:
            violation = violation ** 2
        return violation.mean(), value, violation


class GT(_Comparator):
    """``left >= right``."""
    tag = "gt"

    def forward(self, left, right):
        left, right = _same_device(left, right)
        value = right - left
        violation = F.relu(value)
        if self.norm == 2:
            violation = violation ** 2
        return violation.mean(), value, violation


class Eq(_Comparator):
    """``left == right``."""
    tag = "eq"

    def forward(self, left, right):
        left, right = _same_device(left, right)
        value = left - right
        if self.norm == 1:
            return F.l1_loss(*torch.broadcast_tensors(left, right)), value, value.abs()
        return F.mse_loss(*torch.broadcast_tensors(left, right)), value, value ** 2


# ------------------------------------------------------------------------------------------
_BINARY = {
    "add": operator.add, "sub": operator.sub, "mul": operator.mul, "div": operator.truediv,
    "floordiv": operator.floordiv, "matmul": operator.matmul, "pow": operator.pow,
    "mod": operator.mod,
}
_UNARY = {"neg": operator.neg, "abs": operator.abs, "T": lambda v: v.T, "mT": lambda v: v.mT}


class Variable(nn.Module):
    """A node of a symbolic expression over the keys of a data dictionary.

    ``Variable(key='x')`` reads ``data['x']``; arithmetic, slicing and torch functions applied
    to variables build new variables; comparison operators build ``Constraint`` objects.
    """

    def __init__(self, input_variables=[], func=None, key=None, display_name=None, value=None,
                 op=None, op_args=None):
        super().__init__()
        self._func = func
        if isinstance(value, torch.Tensor) and value.requires_grad:
            value = nn.Parameter(value)
        self._value = value
        children = [self._wrap(v) for v in input_variables]
        self.children_vars = nn.ModuleList(children)
        self._is_input = key is not None
        if op is None:
            op = "input" if self._is_input else ("call" if func is not None else "const")
        self.op, self.op_args = op, op_args
        self._key = key if key is not None else str(id(self))
        if self._is_input:
            assert key not in {n._key for n in self.ordered_nodes[:-1]}, \
                f"Key {key} repeats existing key. Keys should be unique."
        self._display_name = display_name

    # -- construction helpers ---------------------------------------------------------------
    @staticmethod
    def _wrap(obj):
        if isinstance(obj, Variable):
            return obj
        if isinstance(obj, torch.Tensor):
            return Variable(value=obj, display_name=str(obj))
        return Variable(value=obj, display_name=str(obj), op="const")

    def _binary(self, other, tag, reverse=False):
        args = [other, self] if reverse else [self, other]
        return Variable(input_variables=args, func=_BINARY[tag], op=tag, display_name=tag)

    def _unary(self, tag):
        return Variable(input_variables=[self], func=_UNARY[tag], op=tag, display_name=tag)

    # -- naming -----------------------------------------------------------------------------
    @property
    def key(self):
        return self._key

    @key.setter
    def key(self, k):
        self._key = str(id(self)) if k is None else k

    @property
    def display_name(self):
        return self._display_name if self._display_name is not None else self._key

    @property
    def ordered_nodes(self) -> List["Variable"]:
        """Post-order (children before parents) list of distinct nodes, ending with ``self``."""
        seen, order = set(), []

        def visit(node):
            if id(node) in seen:
                return
            seen.add(id(node))
            for c in node.children_vars:
                visit(c)
            order.append(node)
        visit(self)
        return order

    @property
    def keys(self) -> List[str]:
        return [n._key for n in self.ordered_nodes if n._is_input]

    @property
    def value(self):
        return self._value

    def __hash__(self):
        return id(self)

    def __repr__(self):
        return str(self.display_name)

    # -- evaluation -------------------------------------------------------------------------
    def forward(self, datadict: Dict[str, torch.Tensor] = None):
        datadict = {} if datadict is None else datadict
        memo: Dict[int, object] = {}

        def ev(node):
            got = memo.get(id(node))
            if got is not None or id(node) in memo:
                return got
            if node._is_input:
                out = datadict[node._key]
            elif node._func is not None:
                out = node._func(*[ev(c) for c in node.children_vars])
            else:
                out = node._value
            memo[id(node)] = out
            return out
        result = ev(self)
        if not self._is_input and self._func is not None:
            self._value = result
        return result

    # -- operators --------------------------------------------------------------------------
    def __add__(self, o): return self._binary(o, "add")
    def __radd__(self, o): return self._binary(o, "add", True)
    def __sub__(self, o): return self._binary(o, "sub")
    def __rsub__(self, o): return self._binary(o, "sub", True)
    def __mul__(self, o): return self._binary(o, "mul")
    def __rmul__(self, o): return self._binary(o, "mul", True)
    def __truediv__(self, o): return self._binary(o, "div")
    def __rtruediv__(self, o): return self._binary(o, "div", True)
    def __floordiv__(self, o): return self._binary(o, "floordiv")
    def __rfloordiv__(self, o): return self._binary(o, "floordiv", True)
    def __matmul__(self, o): return self._binary(o, "matmul")
    def __rmatmul__(self, o): return self._binary(o, "matmul", True)
    def __pow__(self, o): return self._binary(o, "pow")
    def __rpow__(self, o): return self._binary(o, "pow", True)
    def __mod__(self, o): return self._binary(o, "mod")
    def __rmod__(self, o): return self._binary(o, "mod", True)
    def __neg__(self): return self._unary("neg")
    def __abs__(self): return self._unary("abs")

    @property
    def T(self): return self._unary("T")

    @property
    def mT(self): return self._unary("mT")

    def __getitem__(self, index):
        return Variable(input_variables=[self], func=lambda v: v[index], op="getitem",
                        op_args=index, display_name="slice")

    def __eq__(self, other): return Constraint(self, other, Eq())
    def __lt__(self, other): return Constraint(self, other, LT())
    def __le__(self, other): return Constraint(self, other, LT())
    def __gt__(self, other): return Constraint(self, other, GT())
    def __ge__(self, other): return Constraint(self, other, GT())

    def unpack(self, spec):
        """Split a tuple-valued variable: ``unpack(3)`` or ``unpack(['u', 's', 'v'])``."""
        if isinstance(spec, int):
            names = [None] * spec
        else:
            names = list(spec)
        return [Variable(input_variables=[self], func=(lambda v, i=i: v[i]), op="getitem",
                         op_args=i, display_name=n) for i, n in enumerate(names)]

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        var_args = [a for a in args if isinstance(a, Variable)]
        other_args = tuple(a for a in args if not isinstance(a, Variable))

        def call(*vals, **kw):
            return func(*(other_args + vals), **{**kwargs, **kw})
        return Variable(input_variables=var_args, func=call, op="call", op_args=func,
                        display_name=getattr(func, "__name__", "torch_function"))

    def grad(self, other):
        def _g(y, x):
            return torch.autograd.grad(y, [x], grad_outputs=torch.ones_like(y), create_graph=True)[0]
        return Variable(input_variables=[self, other], func=_g, op="call",
                        display_name=f"d{self.display_name}/d{other.display_name}")

    def minimize(self, metric=torch.mean, weight=1.0, name=None):
        return Objective(self, metric=metric, weight=weight, name=name)

    def show(self, figname=None):
        raise NotImplementedError("graph drawing is outside the hot-path scope of neuromancer_b200")


def variable(*args, display_name=None, **kwargs) -> Variable:
    """Factory with the reference's overloads (constraint.py:622-695):

    ``variable('x')`` input variable; ``variable()`` trainable scalar; ``variable(3, 4)`` or
    ``variable((3, 4))`` trainable tensor; ``variable(tensor)`` constant / trainable value;
    ``variable([inputs...], func)`` arbitrary function node.
    """
    if "key" in kwargs and not args:
        args = (kwargs.pop("key"),)
    if len(args) == 0:
        return Variable(display_name=display_name, value=torch.randn(1, requires_grad=True))
    first = args[0]
    if isinstance(first, str) and len(args) == 1:
        def _never():
            raise RuntimeError("eval_node should never be called on an input_node")
        return Variable(key=first, func=_never)
    if isinstance(first, torch.Tensor) and len(args) == 1:
        return Variable(display_name=display_name, value=first)
    if len(args) == 2 and callable(args[1]) and not isinstance(first, (int, torch.Tensor)):
        return Variable(input_variables=list(first), func=args[1], display_name=display_name)
    if all(isinstance(a, int) and not isinstance(a, bool) for a in args):
        return Variable(display_name=display_name, value=torch.randn(*args, requires_grad=True))
    if isinstance(first, (tuple, list, torch.Size)) and all(isinstance(a, int) for a in first):
        return Variable(display_name=display_name, value=torch.randn(tuple(first), requires_grad=True))
    raise TypeError(f"variable(): unsupported arguments {args!r}")


# ------------------------------------------------------------------------------------------
class Loss(nn.Module):
    """Loss term given by dictionary keys and a callable (reference constraint.py:24-62)."""

    def __init__(self, input_keys: List[str], loss: Callable[..., torch.Tensor], weight=1.0, name="loss"):
        super().__init__()
        self.name, self.input_keys, self.output_keys = name, input_keys, [name]
        self.weight, self.loss = weight, loss

    def forward(self, variables):
        return {self.name: self.weight * self.loss(*[variables[k] for k in self.input_keys])}


class Objective(nn.Module):
    """``weight * metric(var)`` (reference constraint.py:176-228)."""

    def __init__(self, var, metric=torch.mean, weight=1.0, name=None):
        super().__init__()
        assert type(var) is Variable, f"{var} must be Variable type"
        self.var, self.metric, self.weight = var, metric, weight
        self.name = name if name is not None else f"{var.display_name}_{metric}"
        self.input_keys, self.output_keys = var.keys, [f"{var.key}_{metric}"]

    def forward(self, input_dict):
        return {self.output_keys[0]: self.weight * self.metric(self.var(input_dict))}

    def __mul__(self, w): return Objective(self.var, self.metric, self.weight * w, self.name)
    __rmul__ = __mul__


class Constraint(nn.Module):
    """``weight * comparator(left, right)`` built by ``<``, ``>``, ``==`` on variables; ``^ p``
    selects the norm and ``*`` the weight (reference constraint.py:231-324)."""

    def __init__(self, left, right, comparator, weight=1.0, name=None):
        super().__init__()
        left, right = self._as_variable(left), self._as_variable(right)
        self.left, self.right, self.comparator, self.weight = left, right, comparator, weight
        self.key = f"{left.key}_{comparator}_{right.key}"
        self.name = name if name is not None else f"{left.display_name} {comparator} {right.display_name}"
        self.input_keys = left.keys + right.keys
        self.output_keys = [self.key, f"{self.key}_value", f"{self.key}_violation"]

    @staticmethod
    def _as_variable(side):
        if type(side) is Variable:
            return side
        label = str(side) if isinstance(side, (int, float, complex, bool)) else str(id(side))
        if not isinstance(side, torch.Tensor):
            side = torch.tensor(side)
        return Variable(value=side, display_name=label)

    def update_name(self, name):
        self.name = self.key = name
        self.output_keys = [name, f"{name}_value", f"{name}_violation"]

    @property
    def variable_names(self):
        return [self.left.display_name, self.right.display_name]

    def __xor__(self, norm):
        return Constraint(self.left, self.right, type(self.comparator)(norm=norm),
                          weight=self.weight, name=self.name)

    def __mul__(self, w):
        return Constraint(self.left, self.right, self.comparator, weight=self.weight * w, name=self.name)
    __rmul__ = __mul__

    def __bool__(self):
        return self.left is self.right

    def evaluate(self, input_dict):
        """``(weighted loss, value, violation)`` evaluated eagerly on ``input_dict``."""
        lhs, rhs = self.left(input_dict), self.right(input_dict)
        lhs = lhs if isinstance(lhs, torch.Tensor) else torch.tensor(lhs)
        rhs = rhs if isinstance(rhs, torch.Tensor) else torch.tensor(rhs)
        loss, value, violation = self.comparator(lhs, rhs)
        return self.weight * loss, value, violation

    def forward(self, input_dict):
        return dict(zip(self.output_keys, self.evaluate(input_dict)))

    def grad(self, input_dict, input_key=None):
        out = self.forward(input_dict)[self.key]
        return torch.autograd.grad(out, [input_dict[input_key]], create_graph=True)[0]
