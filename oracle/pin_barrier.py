"""
TEST INFRASTRUCTURE -- golden vectors of ``BarrierLoss`` from the UNMODIFIED reference (loss.py:184-257).

    python -m oracle.pin_barrier           # build container only (needs /root/reference); writes tests/golden/barrier.npz

For every built-in barrier function the reference ``BarrierLoss`` is evaluated on a fixed dictionary (two constraints,
one of them violated on part of the batch, and one objective); inputs and every output the class adds
(``loss``, ``objective_loss``, ``penalty_loss``, ``barrier_loss``, ``<constraint>_barrier``) are stored.
``tests/test_api_mirror.py`` requires ``torch.equal`` between these and ``neuromancer_b200.loss.BarrierLoss``.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402

BARRIERS = ["log10", "log", "inverse", "softexp", "softlog", "expshift"]


def build(lib_constraint, lib_loss, barrier):
    """The same expression with either library's classes (they share the operator API)."""
    x, u = lib_constraint.variable("x"), lib_constraint.variable("u")
    obj = 0.5 * ((u == 0.) ^ 2)
    c_hi = 3.0 * (x < 1.0)
    c_lo = 2.0 * (x > -0.5)
    for t, n in ((obj, "effort"), (c_hi, "x_hi"), (c_lo, "x_lo")):
        t.name = n
    return lib_loss.BarrierLoss([obj], [c_hi, c_lo], barrier=barrier, upper_bound=2.0, shift=0.7, alpha=0.4)


def inputs():
    g = torch.Generator(device="cpu")
    g.manual_seed(1234)
    x = 1.5 * torch.rand(6, 5, 2, generator=g) - 0.4           # crosses both bounds; includes exact hits below
    x[0, 0, 0], x[1, 1, 1] = 1.0, -0.5
    return {"x": x, "u": torch.randn(6, 5, 1, generator=g)}


def main():
    ns = ref_shim.load()
    data = inputs()
    out = {"in_x": data["x"].numpy(), "in_u": data["u"].numpy()}
    for b in BARRIERS:
        loss = build(ns.constraint, ns.loss, b)
        res = loss({k: v.clone() for k, v in data.items()})
        for k in ("loss", "objective_loss", "penalty_loss", "barrier_loss", "x_hi_barrier", "x_lo_barrier"):
            out[f"{b}/{k}"] = torch.as_tensor(res[k]).detach().numpy()
    path = os.path.join(ROOT, "tests", "golden", "barrier.npz")
    np.savez(path, **out)
    print("wrote", path, len(out), "arrays")


if __name__ == "__main__":
    main()
