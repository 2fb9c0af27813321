"""
Dense linear map with the ``slim`` interface.

Only the unstructured map is on the hot path (``slim.Linear`` == ``nn.Linear``,
reference slim/linear.py:104-119).  The parameter names (``linear.weight``, ``weight`` aliasing the
same tensor, likewise ``bias``) are kept so that a reference ``state_dict`` loads unchanged.
"""
import torch
import torch.nn as nn


class LinearBase(nn.Module):
    """Interface every linear map offers: ``in_features``/``out_features``, ``effective_W()``
    (the ``[in, out]`` matrix of ``x @ W``), ``reg_error()``."""

    def __init__(self, insize, outsize, bias=False):
        super().__init__()
        self.in_features, self.out_features = insize, outsize
        self.use_bias = bool(bias)

    def effective_W(self):  # pragma: no cover - abstract
        raise NotImplementedError

    def reg_error(self):
        return torch.zeros((), device=self.effective_W().device)

    def eig(self):
        return torch.linalg.eig(self.effective_W())

    def forward(self, x):
        y = x @ self.effective_W()
        return y + self.bias if self.use_bias else y


class Linear(LinearBase):
    """``y = x W^T + b`` backed by ``torch.nn.Linear`` (weight ``[out, in]``)."""

    def __init__(self, insize, outsize, bias=False, **kwargs):
        super().__init__(insize, outsize, bias=bias)
        self.linear = nn.Linear(insize, outsize, bias=bias)
        # aliases: one Parameter, two names -- matches the reference's state_dict layout
        self.weight = self.linear.weight
        self.bias = self.linear.bias

    def effective_W(self):
        return self.linear.weight.T

    def forward(self, x):
        return self.linear(x)


maps = {"linear": Linear}
