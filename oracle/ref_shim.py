"""
TEST INFRASTRUCTURE -- not part of the product path.

Import shim that loads the *unmodified* pnnl/neuromancer reference from
``/root/reference/src`` in the build container, where several of its import-time
dependencies (pydot, plum, lightning, torchdiffeq, torchsde, matplotlib) are absent.

Only used by ``oracle/make_golden.py`` and ``oracle/pin_oracle.py`` to (a) pin the
oracle restatement bit-exactly against the real reference and (b) generate the golden
fixtures committed under ``tests/golden``.  ``/root/reference`` does not exist on the
GPU box, so nothing in ``tests -m gpu``, ``smoke()`` or ``bench.py`` imports this.

Recipe:
  * pre-seed ``sys.modules['neuromancer']`` with an empty package whose ``__path__``
    points at the reference sources, so ``neuromancer/__init__.py`` (which needs
    package metadata, lightning, pyts) never runs;
  * stub the plotting / lightning / torchdiffeq / torchsde modules (never touched by
    the hot path arithmetic);
  * provide a *functional* ``plum.dispatch`` with most-specific-match semantics, because
    ``constraint.variable`` (reference constraint.py:622-695) is a multi-dispatch factory.
"""
from __future__ import annotations

import inspect
import os
import sys
import types
import typing

REFERENCE_SRC = os.environ.get("NM_REFERENCE_SRC", "/root/reference/src")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "neuromancer"))


def _stub(name: str, **attrs) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


class _Anything:
    """No-op object standing in for pydot graph elements."""

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, item):
        def _noop(*a, **k):
            return None
        return _noop


# --------------------------------------------------------------------------- plum shim
def _matches(value, annotation) -> typing.Tuple[bool, int]:
    """Return (matches, specificity).  Higher specificity wins."""
    if annotation is inspect.Parameter.empty:
        return True, 0
    origin = typing.get_origin(annotation)
    if origin is typing.Union:
        best = -1
        for a in typing.get_args(annotation):
            ok, s = _matches(value, a)
            if ok:
                best = max(best, s)
        return (best >= 0), max(best, 0)
    if annotation is type(None):
        return value is None, 3
    if origin is not None:
        import collections.abc as cabc
        if origin in (cabc.Iterable, typing.Iterable):
            if isinstance(value, (str, bytes)) or not isinstance(value, cabc.Iterable):
                # a str is iterable, but the reference's overloads never mean that
                args = typing.get_args(annotation)
                if args and args[0] is str and isinstance(value, (list, tuple)):
                    return all(isinstance(v, str) for v in value), 2
                return False, 0
            args = typing.get_args(annotation)
            if args and args[0] is not typing.Any:
                try:
                    vals = list(value)
                except TypeError:
                    return False, 0
                for v in vals:
                    ok, _ = _matches(v, args[0])
                    if not ok:
                        return False, 0
            return True, 2
        return isinstance(value, origin), 2
    if annotation is typing.Callable or annotation is getattr(__import__("collections.abc").abc, "Callable"):
        return callable(value), 2
    if isinstance(annotation, type):
        if annotation is int and isinstance(value, bool):
            return False, 0
        return isinstance(value, annotation), 3
    return True, 0


class _Dispatcher:
    def __init__(self, name):
        self.name = name
        self.impls = []

    def register(self, fn):
        self.impls.append((fn, inspect.signature(fn)))

    def resolve(self, args, kwargs):
        best, best_score = None, None
        for fn, sig in self.impls:
            try:
                bound = sig.bind(*args, **kwargs)
            except TypeError:
                continue
            score, ok = 0, True
            for pname, val in bound.arguments.items():
                p = sig.parameters[pname]
                ann = p.annotation
                if p.kind is inspect.Parameter.VAR_POSITIONAL:
                    for v in val:
                        m, s = _matches(v, ann)
                        ok &= m
                        score += s
                elif p.kind is inspect.Parameter.VAR_KEYWORD:
                    continue
                else:
                    m, s = _matches(val, ann)
                    ok &= m
                    score += s
            if not ok:
                continue
            if best is None or score > best_score:
                best, best_score = fn, score
        if best is None:
            raise TypeError(f"no overload of {self.name} matches {args} {kwargs}")
        return best


_registry: dict = {}


def _dispatch(fn):
    key = (fn.__module__, fn.__qualname__)
    disp = _registry.get(key)
    if disp is None:
        disp = _registry[key] = _Dispatcher(fn.__qualname__)
    disp.register(fn)

    def wrapper(*args, **kwargs):
        return disp.resolve(args, kwargs)(*args, **kwargs)

    wrapper.__name__ = fn.__name__
    wrapper.__qualname__ = fn.__qualname__
    wrapper.__doc__ = fn.__doc__
    return wrapper


_installed = False


def install() -> None:
    """Make ``import neuromancer.system`` etc. resolve to the reference sources."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError(f"reference sources not found under {REFERENCE_SRC}")
    import torch.nn as nn

    if "pydot" not in sys.modules:
        _stub("pydot", Dot=_Anything, Node=_Anything, Edge=_Anything, Cluster=_Anything)
    if "matplotlib" not in sys.modules:
        mpl = _stub("matplotlib")
        mpl.image = _stub("matplotlib.image", imread=lambda *a, **k: None)
        mpl.pyplot = _stub("matplotlib.pyplot", figure=lambda *a, **k: None,
                           show=lambda *a, **k: None, imshow=lambda *a, **k: None)
    if "lightning" not in sys.modules:
        class LightningModule(nn.Module):
            pass

        class LightningDataModule:
            pass

        lit = _stub("lightning")
        pl = _stub("lightning.pytorch", LightningModule=LightningModule,
                   LightningDataModule=LightningDataModule, Trainer=_Anything)
        lit.pytorch = pl
        cbs = _stub("lightning.pytorch.callbacks", ModelCheckpoint=_Anything)
        es = _stub("lightning.pytorch.callbacks.early_stopping", EarlyStopping=_Anything)
        pl.callbacks = cbs
        cbs.early_stopping = es
    if "torchdiffeq" not in sys.modules:
        tde = _stub("torchdiffeq", odeint_adjoint=None, odeint=None)
        impl = _stub("torchdiffeq._impl")
        adj = _stub("torchdiffeq._impl.adjoint", find_parameters=lambda m: list(m.parameters()))
        tde._impl = impl
        impl.adjoint = adj
    if "torchsde" not in sys.modules:
        _stub("torchsde", sdeint=None, sdeint_adjoint=None)
    if "plum" not in sys.modules:
        _stub("plum", dispatch=_dispatch)

    if "neuromancer" not in sys.modules:
        pkg = types.ModuleType("neuromancer")
        pkg.__path__ = [os.path.join(REFERENCE_SRC, "neuromancer")]
        sys.modules["neuromancer"] = pkg
        # slim's __init__ imports butterfly (pure python fallback) -- fine.
    _installed = True


def load():
    """Import the hot-path modules of the reference; returns a namespace."""
    install()
    import importlib
    ns = types.SimpleNamespace()
    ns.system = importlib.import_module("neuromancer.system")
    ns.constraint = importlib.import_module("neuromancer.constraint")
    ns.loss = importlib.import_module("neuromancer.loss")
    ns.problem = importlib.import_module("neuromancer.problem")
    ns.integrators = importlib.import_module("neuromancer.dynamics.integrators")
    ns.ode = importlib.import_module("neuromancer.dynamics.ode")
    ns.blocks = importlib.import_module("neuromancer.modules.blocks")
    ns.activations = importlib.import_module("neuromancer.modules.activations")
    ns.functions = importlib.import_module("neuromancer.modules.functions")
    ns.slim = importlib.import_module("neuromancer.slim")
    return ns
