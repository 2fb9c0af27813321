"""Parity of the exogenous first-layer hoist (nm_kernels.cuh: HoistedMlp, nm_exo_gemm_*_kernel).

The hoist was written at the end of round 1 without GPU access and is compiled in only with NM_TUNE=hoist=1, so this
test is opt-in (NM_TEST_HOIST=1): it builds a one-class library of c4 / c5 with the flag, loads it in a fresh process
through NM_LIB_PATH and compares the product path with the CPU oracle at the tolerances of test_parity_gpu.py
(with and without the checkpoint buffer: `system(...)` under no_grad passes none, so the adjoint-side recompute of E
is exercised by the raw step)."""
import os
import subprocess
import sys

import pytest

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("NM_TEST_HOIST") != "1", reason="opt-in: set NM_TEST_HOIST=1 (builds a variant library)")]

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import sys, torch
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
import helpers as H
import ctypes as C
from neuromancer_b200 import _lib
name = {name!r}
info = _lib.load().nm_kernel_info(C.byref(H.configs.build_case(name).plan().c)).decode()
print(info)
for B in (1, 127, 300):
    case = H.make_case(name, "cuda:0", seed=3)
    batch_cpu = case.make_batch(B, "cpu", 7)
    params_cpu = [p.detach().cpu().clone() for p in case.plan_params]
    batch = {{k: (v.to("cuda:0") if isinstance(v, torch.Tensor) else v) for k, v in batch_cpu.items()}}
    got = H.product_run(case, batch)
    ref = H.oracle_run(name, params_cpu, batch_cpu)
    for k, v in ref["traj"].items():
        assert H.rel_err(got["traj"][k], v) <= 1e-5, (name, B, k, H.rel_err(got["traj"][k], v))
    assert H.rel_err(got["loss"], ref["loss"].detach()) <= 1e-4, (name, B)
    for i, (a, b) in enumerate(zip(got["grads"], ref["grads"])):
        assert H.rel_err(a, b) <= 1e-4, (name, B, i, H.rel_err(a, b))
# the tensor-core adjoint of the hoisted class (c4: policy mode + E pre-activation + dZ_1 store), forced at a small batch
import os
if "bwd[tcgen05" in info:
    os.environ["NM_TC_BWD"] = "1"
    case = H.make_case(name, "cuda:0", seed=3)
    batch_cpu = case.make_batch(300, "cpu", 9)
    params_cpu = [p.detach().cpu().clone() for p in case.plan_params]
    batch = {{k: (v.to("cuda:0") if isinstance(v, torch.Tensor) else v) for k, v in batch_cpu.items()}}
    got = H.product_run(case, batch)
    ref = H.oracle_run(name, params_cpu, batch_cpu)
    for i, (a, b) in enumerate(zip(got["grads"], ref["grads"])):
        assert H.rel_err(a, b) <= 1e-4, (name, "tensor-core adjoint", i, H.rel_err(a, b))
print("hoist parity ok", name)
"""


@pytest.mark.parametrize("name", ["c4", "c5"])
def test_hoisted_first_layer_matches_oracle(name, tmp_path):
    from neuromancer_b200 import build
    lib = str(tmp_path / f"lib_{name}_hoist.so")
    build.build_variant(name, "hoist=1", lib)
    env = dict(os.environ, NM_LIB_PATH=lib)
    res = subprocess.run([sys.executable, "-c", CHILD.format(root=ROOT, name=name)], env=env, stdout=subprocess.PIPE,
                         stderr=subprocess.STDOUT, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-4000:]
    assert "hoist parity ok" in res.stdout
