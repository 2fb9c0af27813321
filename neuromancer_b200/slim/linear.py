"""
Linear maps with the ``slim`` interface (reference slim/linear.py).

``Linear`` (== ``nn.Linear``, reference slim/linear.py:104-119) is the dense map of the benchmark configurations.
The STRUCTURED maps below parametrise their ``[in, out]`` matrix through ``effective_W()`` (reference
slim/linear.py:85-101); the fused rollout supports them by materialising ``effective_W()`` once per call into the
kernels' flat parameter buffer: the kernels differentiate w.r.t. the materialised matrix and torch's
autograd carries that gradient through ``effective_W()`` to the map's own parameters.  Any module exposing
``effective_W()`` / ``in_features`` / ``out_features`` / ``bias`` is accepted the same way (duck typing: the
reference's own classes included, see ``neuromancer_b200.adopt``).  Parameter names and shapes follow the reference
so that its ``state_dict`` loads unchanged (bias of a structured map: ``[1, out]``).

Not mirrored (outside the hot path; usable through duck typing if imported from the reference): L0Linear, LassoLinear
(rewrites its parameters inside ``forward``), ButterflyLinear, OrthogonalLinear, SchurDecompositionLinear,
SpectralLinear, SymplecticLinear, GershgorinLinear, BoundedNormLinear, TrivialNullSpaceLinear, PowerBoundLinear.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


class LinearBase(nn.Module):
    """Interface every linear map offers: ``in_features``/``out_features``, ``effective_W()`` (the ``[in, out]`` matrix
    of ``x @ W``), ``reg_error()``, ``eig()`` (reference slim/linear.py:37-101)."""

    def __init__(self, insize, outsize, bias=False, provide_weights=True):
        super().__init__()
        self.in_features, self.out_features = insize, outsize
        self.use_bias = bool(bias)
        if bias:
            lim = 1.0 / math.sqrt(insize)
            self.bias = nn.Parameter(torch.empty(1, outsize).uniform_(-lim, lim))
        else:
            self.register_parameter("bias", None)
        if provide_weights:
            w = torch.empty(insize, outsize)
            nn.init.kaiming_uniform_(w, a=math.sqrt(5))
            self.weight = nn.Parameter(w)

    @property
    def device(self):
        return next(self.parameters()).device

    def effective_W(self):  # pragma: no cover - abstract
        raise NotImplementedError

    def reg_error(self):
        return torch.tensor(0.0, device=self.device)

    def eig(self):
        return torch.linalg.eig(self.effective_W())

    def forward(self, x):
        y = torch.matmul(x, self.effective_W())
        return y + self.bias if self.use_bias else y


class Linear(LinearBase):
    """``y = x W^T + b`` backed by ``torch.nn.Linear`` (weight ``[out, in]``)."""

    def __init__(self, insize, outsize, bias=False, **kwargs):
        super().__init__(insize, outsize, bias=bias, provide_weights=False)     # (registers the `bias` slot first, like the reference)
        self.linear = nn.Linear(insize, outsize, bias=bias)
        # aliases: one Parameter, two names -- matches the reference's state_dict layout
        self.weight = self.linear.weight
        self.bias = self.linear.bias

    def effective_W(self):
        return self.linear.weight.T

    def forward(self, x):
        return self.linear(x)


class SquareLinear(LinearBase):
    """Maps that only exist for ``insize == outsize`` (reference slim/linear.py:211-221)."""

    def __init__(self, insize, outsize, bias=False, provide_weights=True, **kwargs):
        if insize != outsize:
            raise AssertionError(f"Map must be square. insize={insize} and outsize={outsize}")
        super().__init__(insize, outsize, bias=bias, provide_weights=provide_weights)


class IdentityInitLinear(Linear):
    """Dense map initialised to the identity (reference slim/linear.py:224-235)."""

    def __init__(self, insize, outsize, bias=False, **kwargs):
        super().__init__(insize, outsize, bias=bias)
        nn.init.eye_(self.linear.weight)
        if bias:
            nn.init.zeros_(self.linear.bias)


class IdentityLinear(IdentityInitLinear):
    """Frozen identity (reference slim/linear.py:238-244)."""

    def __init__(self, insize, outsize, bias=False, **kwargs):
        super().__init__(insize, outsize, bias=bias)
        self.linear.requires_grad_(False)


class NonNegativeLinear(LinearBase):
    """``W = relu(weight)`` (reference slim/linear.py:247-256)."""

    def __init__(self, insize, outsize, bias=False, **kwargs):
        super().__init__(insize, outsize, bias=bias)
        self.weight = nn.Parameter(self.weight.detach().abs() * 0.1)

    def effective_W(self):
        return F.relu(self.weight)


class PSDLinear(SquareLinear):
    """``W = weight^T weight`` (reference slim/linear.py:259-267)."""

    def effective_W(self):
        return self.weight.T @ self.weight


class LassoLinearRELU(LinearBase):
    """``W = relu(u) - relu(v)`` with an L1 shrinkage penalty (reference slim/linear.py:299-320)."""

    def __init__(self, insize, outsize, bias=False, gamma=1.0, **kwargs):
        super().__init__(insize, outsize, bias=bias, provide_weights=False)
        for name in ("u_param", "v_param"):
            t = torch.empty(insize, outsize)
            nn.init.kaiming_normal_(t)
            setattr(self, name, nn.Parameter(t.abs() / 2.0))
        self.gamma = gamma

    def effective_W(self):
        return F.relu(self.u_param) - F.relu(self.v_param)

    def reg_error(self):
        return self.gamma * self.effective_W().norm(p=1)


class RightStochasticLinear(LinearBase):
    """Rows sum to one: ``softmax(weight, dim=1)`` (reference slim/linear.py:352-363)."""

    def effective_W(self):
        return F.softmax(self.weight, dim=1)


class LeftStochasticLinear(LinearBase):
    """Columns sum to one: ``softmax(weight, dim=0)`` (reference slim/linear.py:366-377)."""

    def effective_W(self):
        return F.softmax(self.weight, dim=0)


class PerronFrobeniusLinear(LinearBase):
    """Row sums between ``sigma_min`` and ``sigma_max`` (reference slim/linear.py:380-403)."""

    def __init__(self, insize, outsize, bias=False, sigma_min=0.8, sigma_max=1.0, **kwargs):
        super().__init__(insize, outsize, bias=bias)
        self.scaling = nn.Parameter(torch.rand(insize, outsize))
        self.sigma_min, self.sigma_max = sigma_min, sigma_max

    def effective_W(self):
        row_scale = self.sigma_max - (self.sigma_max - self.sigma_min) * torch.sigmoid(self.scaling)
        return row_scale * F.softmax(self.weight, dim=1)


class SymmetricLinear(SquareLinear):
    """``W = (weight + weight^T) / 2`` (reference slim/linear.py:406-417)."""

    def effective_W(self):
        return (self.weight + self.weight.T) / 2


class SkewSymmetricLinear(SquareLinear):
    """``W = triu(weight) - triu(weight)^T`` (reference slim/linear.py:420-432)."""

    def effective_W(self):
        upper = self.weight.triu()
        return upper - upper.T


class DampedSkewSymmetricLinear(SkewSymmetricLinear):
    """Skew-symmetric part plus ``gamma I`` (reference slim/linear.py:435-448)."""

    def __init__(self, insize, outsize, bias=False, sigma_min=0.1, sigma_max=0.5, **kwargs):
        super().__init__(insize, outsize, bias=bias)
        self.eye = nn.Parameter(torch.eye(insize, outsize), requires_grad=False)
        self.gamma = nn.Parameter(sigma_min + (sigma_max - sigma_min) * torch.rand(1, 1))

    def effective_W(self):
        return super().effective_W() + self.gamma * self.eye


class SplitLinear(LinearBase):
    """``W = B - C`` with non-negative ``B``, ``C`` (reference slim/linear.py:451-465)."""

    def __init__(self, insize, outsize, bias=False, **kwargs):
        super().__init__(insize, outsize, bias=bias, provide_weights=False)
        self.B = NonNegativeLinear(insize, outsize, bias)
        self.C = NonNegativeLinear(insize, outsize, bias)

    def effective_W(self):
        return self.B.effective_W() - self.C.effective_W()


class StableSplitLinear(LinearBase):
    """``W = B - C`` with Perron-Frobenius bounded ``B``, ``C`` (reference slim/linear.py:468-482)."""

    def __init__(self, insize, outsize, bias=False, sigma_min=0.1, sigma_max=1.0, **kwargs):
        super().__init__(insize, outsize, bias=bias, provide_weights=False)
        self.B = PerronFrobeniusLinear(insize, outsize, bias, sigma_max, sigma_max)
        self.C = PerronFrobeniusLinear(insize, outsize, bias, 0, sigma_max - sigma_min)

    def effective_W(self):
        return self.B.effective_W() - self.C.effective_W()


class SVDLinear(LinearBase):
    """``W = U diag(s) V`` with singular values squashed into ``[sigma_min, sigma_max]`` and a soft orthogonality penalty
    on the factors (reference slim/linear.py:485-541)."""

    def __init__(self, insize, outsize, bias=False, sigma_min=0.1, sigma_max=1.0, **kwargs):
        super().__init__(insize, outsize, bias=bias, provide_weights=False)
        self.U = nn.Parameter(nn.init.orthogonal_(torch.empty(insize, insize)))
        self.V = nn.Parameter(nn.init.orthogonal_(torch.empty(outsize, outsize)))
        self.sigma = nn.Parameter(torch.rand(insize, 1))
        self.sigma_min, self.sigma_max = sigma_min, sigma_max

    @staticmethod
    def orthogonal_error(m):
        eye = torch.eye(m.shape[0], device=m.device)
        return torch.norm(torch.norm(eye - m @ m.T, 2) + torch.norm(eye - m.T @ m, 2), 2)

    def reg_error(self):
        return self.orthogonal_error(self.U) + self.orthogonal_error(self.V)

    def effective_W(self):
        s = self.sigma_max - (self.sigma_max - self.sigma_min) * torch.sigmoid(self.sigma)
        core = torch.eye(self.in_features, self.out_features, device=self.sigma.device) * s
        return self.U @ (core @ self.V)


class SVDLinearLearnBounds(SVDLinear):
    """``SVDLinear`` with trainable singular-value bounds (reference slim/linear.py:544-551)."""

    def __init__(self, insize, outsize, bias=False, sigma_min=0.1, sigma_max=1.0, **kwargs):
        super().__init__(insize, outsize, bias=bias, sigma_min=sigma_min, sigma_max=sigma_max)
        self.sigma_min = nn.Parameter(torch.tensor(sigma_min))
        self.sigma_max = nn.Parameter(torch.tensor(sigma_max))


class SymmetricSVDLinear(SVDLinear):
    """``U = V`` (reference slim/linear.py:554-560)."""

    def __init__(self, insize, outsize, bias=False, sigma_min=0.1, sigma_max=1.0, **kwargs):
        super().__init__(insize, outsize, bias=bias, sigma_min=sigma_min, sigma_max=sigma_max)
        self.U = self.V


class DerivedWeight:
    """The ``[out, in]`` weight of a structured map as the fused path sees it: not a Parameter but a function of the
    map's parameters, re-evaluated (differentiably) on every rollout call.  Quacks like a tensor where the plan needs it
    (``shape``, ``numel()``, ``requires_grad``)."""

    def __init__(self, module):
        self.module = module
        self.shape = torch.Size((int(module.out_features), int(module.in_features)))

    def numel(self):
        return self.shape[0] * self.shape[1]

    def own_parameters(self):
        """the module Parameters this entry is a function of (the bias is a plan entry of its own)"""
        return [t for n, t in self.module.named_parameters() if n != "bias"]

    @property
    def requires_grad(self):
        return any(p.requires_grad for p in self.own_parameters())

    def materialise(self):
        w = self.module.effective_W()
        if tuple(w.shape) != (self.shape[1], self.shape[0]):
            raise ValueError(f"{type(self.module).__name__}.effective_W() returned {tuple(w.shape)}, expected "
                             f"{(self.shape[1], self.shape[0])}")
        return w.T.contiguous()


maps = {"linear": Linear, "nneg": NonNegativeLinear, "lasso_relu": LassoLinearRELU, "lstochastic": LeftStochasticLinear,
        "rstochastic": RightStochasticLinear, "pf": PerronFrobeniusLinear, "symmetric": SymmetricLinear,
        "skew_symetric": SkewSymmetricLinear, "damp_skew_symmetric": DampedSkewSymmetricLinear, "split": SplitLinear,
        "stable_split": StableSplitLinear, "softSVD": SVDLinear, "learnSVD": SVDLinearLearnBounds, "psd": PSDLinear,
        "identity": IdentityLinear}
square_maps = {SymmetricLinear, SkewSymmetricLinear, DampedSkewSymmetricLinear, PSDLinear, SymmetricSVDLinear}
