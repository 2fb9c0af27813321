"""
Function-approximator blocks on the hot path: ``Block``, ``Linear``, ``MLP``, ``MLP_bounds``.

API mirror of the reference (``modules/blocks.py``: ``Block`` :20-43, ``MLP`` :126-177,
``MLP_bounds`` :555-619): same constructor signatures, ``in_features``/``out_features``,
``linear``/``nonlin`` ``ModuleList`` names (so ``state_dict`` keys are unchanged).

Inside a ``System`` these blocks are never executed layer by layer: the plan compiler
(``neuromancer_b200.plan``) reads their weights and lowers them to an ``NmMlp`` descriptor that the
fused sm_100a kernels evaluate.  ``forward`` below is the single-module eager call for use outside
a rollout (e.g. inspecting a trained policy on one state).
"""
import torch
import torch.nn as nn

from .. import slim
from .activations import SoftExponential, lower_activation
from .functions import bounds_clamp, bounds_scaling


class Block(nn.Module):
    """Blocks take one or several ``[batch, d_k]`` tensors; several are concatenated on the last
    axis before evaluation (reference blocks.py:32-43)."""

    def block_eval(self, x):  # pragma: no cover - abstract
        raise NotImplementedError

    def forward(self, *inputs):
        x = inputs[0] if len(inputs) == 1 else torch.cat(inputs, dim=-1)
        return self.block_eval(x)


class Linear(Block):
    """A single linear map with the block interface (reference blocks.py:47-86)."""

    def __init__(self, insize, outsize, bias=True, linear_map=slim.Linear, nonlin=None,
                 hsizes=None, linargs=None):
        super().__init__()
        self.in_features, self.out_features = insize, outsize
        self.linear = linear_map(insize, outsize, bias=bias, **(linargs or {}))

    def reg_error(self):
        return self.linear.reg_error()

    def block_eval(self, x):
        return self.linear(x)


def _dense_weight_bias(layer):
    """Return ``(weight [out,in], bias [out] or None)`` of a dense layer, whichever class built it."""
    if isinstance(layer, nn.Linear):
        return layer.weight, layer.bias
    inner = getattr(layer, "linear", None)
    if isinstance(inner, nn.Linear):
        return inner.weight, inner.bias
    if callable(getattr(layer, "effective_W", None)) and hasattr(layer, "in_features") and hasattr(layer, "out_features"):
        # structured slim map (ours or the reference's): the fused path sees the materialised effective_W()
        return slim.DerivedWeight(layer), getattr(layer, "bias", None)
    raise NotImplementedError(
        f"linear map {type(layer).__name__} is neither a dense nn.Linear nor a slim map with effective_W()")


class MLP(Block):
    """Fully connected network ``x <- nonlin_k(x W_k^T + b_k)`` with an identity after the last
    layer (reference blocks.py:131-177)."""

    def __init__(self, insize, outsize, bias=True, linear_map=slim.Linear,
                 nonlin=SoftExponential, hsizes=[64], linargs=dict()):
        super().__init__()
        self.in_features, self.out_features = insize, outsize
        hsizes = list(hsizes)
        self.nhidden = len(hsizes)
        widths = [insize, *hsizes, outsize]
        self.nonlin = nn.ModuleList([nonlin() for _ in hsizes] + [nn.Identity()])
        self.linear = nn.ModuleList(
            linear_map(widths[i], widths[i + 1], bias=bias, **linargs)
            for i in range(len(widths) - 1))

    def reg_error(self):
        return sum(m.reg_error() for m in self.linear if hasattr(m, "reg_error"))

    # -- lowering helpers used by neuromancer_b200.plan -------------------------------------
    def dense_layers(self):
        """``[(weight, bias), ...]`` in evaluation order."""
        return [_dense_weight_bias(m) for m in self.linear]

    def layer_sizes(self):
        pairs = self.dense_layers()
        return [pairs[0][0].shape[1]] + [w.shape[0] for w, _ in pairs]

    def lowered_activation(self):
        """``(NM_ACT_*, param)`` shared by all hidden layers; raises if they differ."""
        ids = {lower_activation(m) for m in list(self.nonlin)[:-1]}
        if not isinstance(self.nonlin[-1], nn.Identity):
            raise NotImplementedError("the last nonlinearity of an MLP must be the identity")
        if len(ids) > 1:
            raise NotImplementedError("mixed hidden activations are not fused")
        return ids.pop() if ids else (0, 0.0)

    def bound_spec(self):
        """``(NM_BOUNDS_*, min, max)``; plain MLPs have no output bound."""
        return 0, None, None

    # -- eager single-call evaluation -------------------------------------------------------
    def _hidden_chain(self, x):
        for lin, act in zip(self.linear, self.nonlin):
            x = act(lin(x))
        return x

    def block_eval(self, x):
        return self._hidden_chain(x)


class MLP_bounds(MLP):
    """``MLP`` whose output is mapped into ``[min, max]`` by ``sigmoid_scale`` or ``relu_clamp``
    (reference blocks.py:555-619)."""

    bound_methods = {"sigmoid_scale": bounds_scaling, "relu_clamp": bounds_clamp}

    def __init__(self, insize, outsize, bias=True, linear_map=slim.Linear,
                 nonlin=SoftExponential, hsizes=[64], linargs=dict(),
                 min=0.0, max=1.0, method="sigmoid_scale"):
        super().__init__(insize=insize, outsize=outsize, bias=bias, linear_map=linear_map,
                         nonlin=nonlin, hsizes=hsizes, linargs=linargs)
        self.min, self.max = min, max
        self.method = self._set_method(method)

    def _set_method(self, method):
        if method in self.bound_methods:
            return self.bound_methods[method]
        assert callable(method), (
            f"Method, {method} must be a key in {self.bound_methods} or a differentiable callable.")
        return method

    def bound_spec(self):
        if self.method is bounds_scaling:
            kind = 1
        elif self.method is bounds_clamp:
            kind = 2
        else:
            raise NotImplementedError("custom bound callables are not fused")
        return kind, self.min, self.max

    def block_eval(self, x):
        return self.method(self._hidden_chain(x), self.min, self.max)


blocks = {"mlp": MLP, "mlp_bounds": MLP_bounds, "linear": Linear}
