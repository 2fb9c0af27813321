"""The oracle restatement against the golden vectors produced by the UNMODIFIED reference
(oracle/pin_oracle.py ran the real pnnl/neuromancer through oracle/ref_shim.py and recorded them).
CPU only."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import cases as ocases
from oracle import dpc_oracle

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
NAMES = sorted(ocases.CASES)


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, f"{name}.npz"), allow_pickle=False)
    params = [torch.from_numpy(z[f"param_{i}"]) for i in range(sum(k.startswith("param_") for k in z.files))]
    inputs = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("in_")}
    traj = {k[5:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("traj_") and k != "traj_keys"}
    terms = {str(n): torch.from_numpy(z[f"term_{n}"]) for n in z["term_names"]}
    grads = [torch.from_numpy(z[f"grad_{i}"]) for i in range(len(params))]
    scal = {k: torch.from_numpy(z[k]) for k in ("objective_loss", "penalty_loss", "loss")}
    cmat = {k[5:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("cmat_")}
    return dict(params=params, inputs=inputs, traj=traj, terms=terms, grads=grads, cmat=cmat, **scal)


def test_pin_report_says_bit_exact():
    rep = json.load(open(os.path.join(GOLDEN, "PINNED.json")))
    assert rep["all_bit_exact"], rep
    assert set(rep["cases"]) == set(NAMES)


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_reference_golden(name):
    g = load_golden(name)
    params = [p.clone().requires_grad_(True) for p in g["params"]]
    res = dpc_oracle.run_case(ocases.build(name, params), g["inputs"])
    # same torch build + CPU => identical; a different host CPU may pick another SIMD path, so allow a few ulp
    tol = dict(rtol=2e-6, atol=1e-7)
    for k, v in g["traj"].items():
        torch.testing.assert_close(res["traj"][k], v, **tol)
    for k, v in g["terms"].items():
        torch.testing.assert_close(res["terms"][k].detach(), v, rtol=2e-6, atol=0)
    torch.testing.assert_close(torch.as_tensor(res["loss"]).detach(), g["loss"], rtol=2e-6, atol=0)
    for a, b in zip(res["grads"], g["grads"]):
        torch.testing.assert_close(a, b, rtol=5e-5, atol=1e-6 * float(b.abs().max()))


def test_oracle_fp64_agrees_with_fp32_golden():
    """fp64 run of the oracle: the fp32 reference numbers are within fp32 round-off of it."""
    g = load_golden("c2")
    params = [p.double().clone().requires_grad_(True) for p in g["params"]]
    inputs = {k: v.double() for k, v in g["inputs"].items()}
    res = dpc_oracle.run_case(ocases.build("c2", params), inputs)
    torch.testing.assert_close(res["traj"]["x"].float(), g["traj"]["x"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(torch.as_tensor(res["loss"]).float().detach(), g["loss"], rtol=1e-5, atol=0)


def test_system_forward_append_semantics():
    """reference tests/test_system.py:478-489 pins the slice/append order: a node sees the same-step
    output of an earlier node but the pre-update state."""
    nodes = [dict(fn=lambda x: 2 * x, inputs=["x"], outputs=["u"]),
             dict(fn=lambda x, u: x + u, inputs=["x", "u"], outputs=["x"])]
    out = dpc_oracle.system_forward(nodes, {"x": torch.ones(3, 1, 2)}, nsteps=3)
    assert out["x"].shape == (3, 4, 2) and out["u"].shape == (3, 3, 2)
    assert torch.equal(out["x"][0, :, 0], torch.tensor([1., 3., 9., 27.]))
    assert torch.equal(out["u"][0, :, 0], torch.tensor([2., 6., 18.]))
    # nsteps = 0 leaves the data untouched (reference test grid includes (0, 0))
    assert dpc_oracle.system_forward(nodes, {"x": torch.ones(0, 1, 2)}, nsteps=0)["x"].shape == (0, 1, 2)
