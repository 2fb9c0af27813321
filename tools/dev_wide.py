"""Development check of the wide-policy tensor-core kernels (nm_wide_kernels.cuh) on a GPU box:
    python tools/dev_wide.py [B ...]
runs c5 through the product path (tcgen05 forward, then the FFMA adjoint and the forced tensor-core adjoint) and
compares with the CPU oracle.  Test infrastructure (imports oracle/ through tests/helpers.py)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import helpers as H  # noqa: E402
import ctypes as C  # noqa: E402
from neuromancer_b200 import _lib  # noqa: E402


def slice_err(a, b):
    """largest per-time-step relative error: max_t ( max|a-b| over the slice / max|b| over the slice )"""
    a, b = a.double(), b.double()
    d = (a - b).abs().amax(dim=(0, 2))
    s = b.abs().amax(dim=(0, 2)).clamp_min(1e-30)
    return float((d / s).max())


def main():
    name = os.environ.get("NM_DEV_CASE", "c5")
    nsteps = os.environ.get("NM_DEV_NSTEPS")
    kw = {"nsteps": int(nsteps)} if nsteps else {}
    sizes = [int(v) for v in sys.argv[1:]] or [300]
    info = _lib.load().nm_kernel_info(C.byref(H.configs.build_case(name, **kw).plan().c)).decode()
    print(info, flush=True)
    for B in sizes:
        for force in ("0", "1"):
            os.environ["NM_TC_BWD"] = force
            case = H.make_case(name, "cuda:0", seed=3, **kw)
            batch_cpu = case.make_batch(B, "cpu", 7)
            params_cpu = [p.detach().cpu().clone() for p in case.plan_params]
            batch = {k: (v.to("cuda:0") if isinstance(v, torch.Tensor) else v) for k, v in batch_cpu.items()}
            t0 = time.time()
            got = H.product_run(case, batch)
            torch.cuda.synchronize()
            t1 = time.time()
            ref = H.oracle_run(name, params_cpu, batch_cpu, **kw)
            line = [f"B={B} NM_TC_BWD={force} ({t1 - t0:.2f}s gpu, {time.time() - t1:.1f}s oracle)"]
            for k, v in ref["traj"].items():
                line.append(f"traj {k}: max-norm {H.rel_err(got['traj'][k], v):.2e} per-step {slice_err(got['traj'][k], v):.2e}")
            line.append(f"loss {float(got['loss']):.7g} vs {float(ref['loss']):.7g} ({H.rel_err(got['loss'], ref['loss'].detach()):.1e})")
            ge = [H.rel_err(a, b) for a, b in zip(got["grads"], ref["grads"])]
            line.append("grads " + " ".join(f"{e:.1e}" for e in ge))
            print(" | ".join(line), flush=True)


if __name__ == "__main__":
    main()
