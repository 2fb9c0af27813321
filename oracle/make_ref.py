"""
TEST / BASELINE INFRASTRUCTURE -- not part of the product path.

Copies the UNMODIFIED reference package (pnnl/neuromancer, ``/root/reference/src/neuromancer``: Python sources only)
into ``oracle/_ref/src/neuromancer`` so that ``bench.py --impl reference`` can time the reference's own CPU
implementation of the hot path (``System.forward`` / ``Node`` / ``MLP_bounds`` / ``RK4`` / ``PenaltyLoss`` /
``Problem`` + autograd) on the GPU box, where ``/root/reference`` does not exist.  ``oracle/_ref/`` is git-ignored (no
reference source enters the history) but travels with the ``gpurun`` snapshot like the built ``.so`` files.  The
modules are loaded through ``oracle/ref_shim.py`` (stubs for the absent pydot / lightning / plum ... imports); the
builders are the ones ``oracle/pin_oracle.py`` pins the oracle with.

    python -m oracle.make_ref           # run in the build container; __graft_entry__.build() calls it
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/src/neuromancer"
DST = os.path.join(HERE, "_ref", "src", "neuromancer")


def main() -> int:
    if not os.path.isdir(SRC):
        print(f"[make_ref] {SRC} not present: nothing to do (the GPU box uses the prebuilt copy)")
        return 0
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    n = 0
    for root, dirs, files in os.walk(SRC):
        dirs[:] = [d for d in dirs if d != "__pycache__"]
        for f in files:
            if not f.endswith(".py"):
                continue
            rel = os.path.relpath(os.path.join(root, f), SRC)
            out = os.path.join(DST, rel)
            os.makedirs(os.path.dirname(out), exist_ok=True)
            shutil.copyfile(os.path.join(root, f), out)
            n += 1
    print(f"[make_ref] {n} reference modules -> {os.path.relpath(DST, os.path.dirname(HERE))}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
