"""
Activation registry.

Same names as the reference registry (``modules/activations.py:230-256``) for the activations the
fused kernels implement, plus the kernel ids (``NM_ACT_*`` of ``include/nm_rollout.h``) the plan
compiler lowers them to.  Learnable exotic activations of the reference (BLU, APLU, PELU, ...) are
outside the hot-path scope.
"""
import torch
import torch.nn as nn

# ids must match include/nm_rollout.h
NM_ACT_IDENTITY, NM_ACT_RELU, NM_ACT_GELU, NM_ACT_LEAKY = 0, 1, 2, 3
NM_ACT_TANH, NM_ACT_ELU, NM_ACT_SIGMOID, NM_ACT_SOFTPLUS = 4, 5, 6, 7


class SoftExponential(nn.Module):
    """Soft exponential activation (the reference's default ``nonlin`` of ``MLP``,
    activations.py:24).  With the initial ``alpha = 0`` it is the identity.  It has a trainable
    shape parameter, so a rollout using it cannot be lowered to the fused kernels."""

    def __init__(self, alpha=0.0, tune_alpha=True):
        super().__init__()
        self.alpha = nn.Parameter(torch.tensor(float(alpha)), requires_grad=tune_alpha)

    def forward(self, x):
        a = self.alpha
        if a == 0.0:
            return x
        if a < 0.0:
            return -torch.log(1.0 - a * (x + a)) / a
        return (torch.exp(a * x) - 1.0) / a + a


activations = {
    "softexp": SoftExponential,
    "relu": nn.ReLU,
    "gelu": nn.GELU,
    "leakyrelu": nn.LeakyReLU,
    "tanh": nn.Tanh,
    "elu": nn.ELU,
    "sigmoid": nn.Sigmoid,
    "softplus": nn.Softplus,
    "identity": nn.Identity,
}


def lower_activation(module):
    """Map an activation module instance to ``(NM_ACT_*, act_param)``.

    Raises ``NotImplementedError`` for anything the sm_100a kernels do not implement -- there is
    no eager fallback inside a rollout."""
    if isinstance(module, nn.Identity):
        return NM_ACT_IDENTITY, 0.0
    if isinstance(module, nn.ReLU):
        return NM_ACT_RELU, 0.0
    if isinstance(module, nn.GELU):
        if getattr(module, "approximate", "none") != "none":
            raise NotImplementedError("only the exact (erf) GELU is implemented in the fused kernels")
        return NM_ACT_GELU, 0.0
    if isinstance(module, nn.LeakyReLU):
        return NM_ACT_LEAKY, float(module.negative_slope)
    if isinstance(module, nn.Tanh):
        return NM_ACT_TANH, 0.0
    if isinstance(module, nn.ELU):
        return NM_ACT_ELU, float(module.alpha)
    if isinstance(module, nn.Sigmoid):
        return NM_ACT_SIGMOID, 0.0
    if isinstance(module, nn.Softplus):
        if module.beta != 1.0 or module.threshold != 20.0:
            raise NotImplementedError("Softplus is fused only with beta=1, threshold=20")
        return NM_ACT_SOFTPLUS, 0.0
    raise NotImplementedError(
        f"activation {type(module).__name__} has no fused sm_100a implementation")
