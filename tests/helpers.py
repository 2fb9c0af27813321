"""Shared helpers of the parity tests: run one named case through the CUDA product path and through
the CPU oracle on identical weights and inputs."""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from neuromancer_b200 import configs  # noqa: E402
from oracle import cases as ocases  # noqa: E402
from oracle import dpc_oracle  # noqa: E402


def case_kwargs(name, nsteps=None):
    return {} if nsteps is None else {"nsteps": nsteps}


def product_run(case, batch, backward=True):
    """Run Problem.forward (+ backward) on CUDA; returns dict of cpu tensors."""
    out = case.problem(batch)
    lay = case.system.layout()
    res = {"traj": {}, "terms": {}}
    for k in (lay.state_key, lay.action_key, lay.output_key):
        if k is not None:
            res["traj"][k] = out[f"train_{k}"].detach().cpu()
    for term in case.problem.loss.terms():
        res["terms"][term.name] = out[f"train_{term.output_keys[0]}"].detach().cpu()
    for k in ("objective_loss", "penalty_loss", "loss"):
        res[k] = out[f"train_{k}"].detach().cpu()
    if backward:
        params = [p for p in case.plan_params if p.requires_grad]
        for p in params:
            p.grad = None
        if params:                                   # simulation-only cases have nothing to differentiate
            out["train_loss"].backward()
        res["grads"] = [p.grad.detach().cpu().clone() for p in params]
    return res


def oracle_run(name, params_cpu, batch_cpu, dtype=torch.float32, backward=True, **kw):
    params = [p.detach().to(dtype).clone().requires_grad_(True) for p in params_cpu]
    if not params:
        params = []
    data = {k: (v.to(dtype) if isinstance(v, torch.Tensor) else v) for k, v in batch_cpu.items()}
    spec = ocases.build(name, params, **kw)
    return dpc_oracle.run_case(spec, data, want_grad=backward)


def make_case(name, device, seed=0, **kw):
    return configs.make_case(name, device, seed=seed, **kw)


def rel_err(a, b):
    a, b = a.detach().double(), b.detach().double()
    denom = b.abs().max().clamp_min(1e-30)
    return float(((a - b).abs().max() / denom).detach())
