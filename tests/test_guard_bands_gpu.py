"""Out-of-bounds WRITE detector for the kernels behind the C ABI (compute-sanitizer is closed on this GPU pool): every
caller-owned output / scratch buffer of ``nm_rollout_forward`` / ``nm_rollout_backward`` -- trajectories, checkpoints,
gradient + scalars, workspace -- is carved out of a larger allocation whose guard bands hold a bit pattern; after the
forward and adjoint kernels ran on ragged batches (a partial last tile, forced tensor-core adjoints, a chunked horizon
sweep for the wide kernels) the bands must be untouched and the results must equal those of the plain allocation."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import helpers as H  # noqa: E402
from neuromancer_b200 import rollout as nr  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GUARD = 1 << 16                       # bytes on each side
PATTERN = 0x5A


class Guarded:
    def __init__(self, like: torch.Tensor):
        nbytes = like.numel() * like.element_size()
        pad = (-nbytes) % 256
        self.raw = torch.full((GUARD + nbytes + pad + GUARD,), PATTERN, dtype=torch.uint8, device=like.device)
        self.nbytes = nbytes
        self.view = self.raw[GUARD:GUARD + nbytes].view(like.dtype).view(like.shape)
        assert self.view.data_ptr() % 256 == 0

    def intact(self):
        lo, hi = self.raw[:GUARD], self.raw[GUARD + self.nbytes:]
        return bool((lo == PATTERN).all()) and bool((hi == PATTERN).all())


CASES = [("c5", 300, "1", {"NM_WIDE_STAGING_GB": "0.003"}), ("c5", 129, "1", {}), ("c5", 77, "0", {}), ("c4", 300, "1", {}),
         ("c4", 131, "0", {}), ("c2", 300, "1", {}), ("c2", 129, "0", {}), ("c3", 200, "1", {}), ("c1", 333, "1", {}),
         ("t_preview_reflect", 150, "0", {}), ("t_ssm_softplus", 257, "0", {}), ("t_int_rkf_node", 200, "1", {})]


@pytest.mark.parametrize("name,B,force,env", CASES)
def test_kernels_write_only_inside_their_buffers(name, B, force, env):
    old = {k: os.environ.get(k) for k in ("NM_TC_BWD", *env)}
    os.environ["NM_TC_BWD"] = force
    os.environ.update(env)
    try:
        kw = {"nsteps": 24} if name == "c5" else {}
        case = H.make_case(name, DEV, seed=1, **kw)
        batch = {k: (v.to(DEV) if isinstance(v, torch.Tensor) else v) for k, v in case.make_batch(B, "cpu", 3).items()}
        plan = case.plan(batch)
        plain = nr.RawTrainStep(plan, batch)
        plain.forward(); plain.backward()
        torch.cuda.synchronize()
        want = {k: getattr(plain, k).clone() for k in ("x_traj", "gbuf") if getattr(plain, k) is not None}
        step = nr.RawTrainStep(plan, batch)
        guards = {}
        for attr in ("x_traj", "u_traj", "y_traj", "chk", "ws", "gbuf"):
            t = getattr(step, attr)
            if t is None:
                continue
            g = Guarded(t)
            if attr == "gbuf":
                g.view.zero_()
            guards[attr] = g
            setattr(step, attr, g.view)
        n = int(plan.c.n_params)
        step.grad, step.scalars = step.gbuf[:n], step.gbuf[n:]
        step.forward(); step.backward()
        torch.cuda.synchronize()
        for attr, g in guards.items():
            assert g.intact(), f"{name} B={B} NM_TC_BWD={force}: a kernel wrote outside `{attr}`"
        assert torch.equal(step.x_traj, want["x_traj"])
        assert torch.equal(step.gbuf, want["gbuf"])          # deterministic, and independent of where the buffers live
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
