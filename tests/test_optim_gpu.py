"""Fused gradient clipping + AdamW over the flat parameter arena (``nm_adamw_step``) against
``torch.nn.utils.clip_grad_norm_`` + ``torch.optim.AdamW`` on the same problem, gradients and hyper-parameters
(reference trainer.py:251-252)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import helpers as H  # noqa: E402
from neuromancer_b200.optim import FusedAdamW  # noqa: E402
from neuromancer_b200.trainer import Trainer  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("name,clip", [("c2", 0.05), ("c2", 0.0), ("c1", 1e3), ("c5", 0.02)])
def test_fused_adamw_follows_torch_adamw(name, clip):
    kw = dict(lr=3e-3, betas=(0.8, 0.95), eps=1e-7, weight_decay=0.02)
    a, b = H.make_case(name, DEV, seed=9), H.make_case(name, DEV, seed=9)
    batch = {k: (v.to(DEV) if isinstance(v, torch.Tensor) else v) for k, v in a.make_batch(192, "cpu", 5).items()}
    ref_opt = torch.optim.AdamW([p for p in a.problem.parameters() if p.requires_grad], foreach=False, **kw)
    opt = FusedAdamW(b.problem, max_grad_norm=clip, **kw)
    for step in range(6):
        out = a.problem(batch)
        ref_opt.zero_grad()
        out["train_loss"].backward()
        if clip > 0:
            want_norm = torch.nn.utils.clip_grad_norm_([p for p in a.problem.parameters() if p.requires_grad], clip)
        ref_opt.step()
        out_b = b.problem(batch)
        opt.zero_grad()
        out_b["train_loss"].backward()
        opt.step()
        if clip > 0:
            assert abs(float(opt.last_grad_norm) - float(want_norm)) <= 2e-5 * float(want_norm)     # fp32 (torch) vs double (kernel) sum of squares
        for (n, p), q in zip(a.problem.named_parameters(), b.problem.parameters()):
            err = float((p - q).abs().max() / p.abs().max().clamp_min(1e-30))
            assert err <= 2e-6 * (step + 1), f"{name} step {step} parameter {n}: {err:.2e}"
    assert opt.step_count == 6


def test_frozen_parameters_are_left_alone_and_strays_are_refused():
    case = H.make_case("c2", DEV, seed=2)
    batch = {k: (v.to(DEV) if isinstance(v, torch.Tensor) else v) for k, v in case.make_batch(64, "cpu", 1).items()}
    frozen = case.plan_params[0]
    frozen.requires_grad_(False)
    before = frozen.detach().clone()
    opt = FusedAdamW(case.problem, lr=1e-2)
    out = case.problem(batch)
    out["train_loss"].backward()
    opt.step()
    assert torch.equal(frozen, before)
    assert not torch.equal(case.plan_params[2].detach(), H.make_case("c2", DEV, seed=2).plan_params[2].detach())
    slim_case = H.make_case("t_slim_pf", DEV)
    sb = {k: (v.to(DEV) if isinstance(v, torch.Tensor) else v) for k, v in slim_case.make_batch(32, "cpu", 1).items()}
    slim_case.problem(sb)["train_loss"].backward()
    with pytest.raises(RuntimeError, match="structured slim maps"):
        FusedAdamW(slim_case.problem).step()


def test_trainer_runs_the_fused_optimiser():
    case = H.make_case("c2", DEV, seed=4)
    data = [case.make_batch(256, "cpu", s) for s in range(3)]
    tr = Trainer(case.problem, data, optimizer=FusedAdamW(case.problem, lr=2e-3), epochs=3, epoch_verbose=0, clip=10.0, device=DEV)
    first = float(case.problem({k: (v.to(DEV) if isinstance(v, torch.Tensor) else v) for k, v in data[0].items()})["train_loss"])
    tr.train()
    last = float(case.problem({k: (v.to(DEV) if isinstance(v, torch.Tensor) else v) for k, v in data[0].items()})["train_loss"])
    assert last < first
