"""
In-tree build of the sm_100a extension.

    python -m neuromancer_b200.build            # libnm_rollout.so with the ahead-of-time shape classes
    python -m neuromancer_b200.build --force

The fused kernels are C++ templates over a *shape class* (state/action widths, MLP layer sizes,
activation, bounds kind, integrator, right-hand side ...); everything that is a *value* (weights,
step size, constants, horizon, batch, loss terms) is a runtime field of ``NmPlan``.  For every shape
class this module generates a small header (``csrc/specs/<hash>.h``) and compiles ``nm_spec_tu.cu``
against it.  The five benchmark configurations and the shapes used by the tests are compiled ahead
of time into ``_C/libnm_rollout.so``; a plan with an unseen shape class is compiled on first use
into ``_C/nm_spec_<hash>.so`` (``ensure_specialisation``) and loaded with ``nm_load_plugin``.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from typing import Dict, List

from . import _lib as L

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SPEC_DIR = os.path.join(CSRC, "specs")
OUT_DIR = L.LIB_DIR
OBJ_DIR = os.path.join(OUT_DIR, "obj")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default",
    "-I", INCLUDE, "-I", CSRC,
]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build the sm_100a extension")


# ------------------------------------------------------------------------------- signatures
def _mlp_sig(m) -> str:
    sizes = "-".join(str(m.sizes[i]) for i in range(m.n_layers + 1))
    return f"[{sizes}]a{m.act}_b{1 if m.has_bias else 0}_o{m.bounds}"


def signature_of(p: L.NmPlan) -> str:
    """Python twin of ``make_signature`` in csrc/nm_api.cu (tests assert they agree)."""
    s = f"v{L.NM_ABI_VERSION}_nx{p.nx}_nu{p.nu}_ny{p.ny}"
    if p.has_policy:
        s += "_pol" + _mlp_sig(p.policy) + "_seg"
        parts = []
        for i in range(p.n_segs):
            g = p.segs[i]
            if g.kind == L.NM_SEG_PREVIEW:
                parts.append(f"p{g.index}.{g.width}.{p.exo[g.index].width}")
                continue
            ch = {L.NM_SEG_STATE: "x", L.NM_SEG_OUTPUT: "y"}.get(g.kind, "e")
            parts.append(f"{ch}{g.index if g.kind == L.NM_SEG_EXO else 0}.{g.width}")
        s += ",".join(parts)
    integ = p.integrator if p.dyn_kind == L.NM_DYN_ODE else 0
    s += f"_out{1 if p.has_output else 0}_dyn{p.dyn_kind}_int{integ}_rhs{p.rhs_kind}"
    if L.rhs_has_mlp(p.rhs_kind):
        s += _mlp_sig(p.rhs_mlp)
    nd = p.exo[p.dist_exo].width if p.dist_exo >= 0 else 0
    s += f"_dist{p.dist_exo}.{nd}_act{p.act_exo}"
    return s


def _atom_sig(a, t) -> str:
    if a.var < 0:
        return "k"
    return f"v{a.var}{'b' if (a.t_len == 1 and t.t_len > 1) else 'r'}{'b' if (a.d_len == 1 and t.d_len > 1) else 'r'}"


def term_signature_of(p: L.NmPlan) -> str:
    """Python twin of ``make_term_signature`` in csrc/nm_api.cu."""
    parts = []
    for k in range(p.n_terms):
        t = p.terms[k]
        s = f"T{t.cmp}{t.norm}{'o' if t.is_objective else 'c'}"
        for tag, side in (("L", t.lhs), ("R", t.rhs)):
            s += f"{tag}{side.op}" + _atom_sig(side.a, t) + (_atom_sig(side.b, t) if side.op != L.NM_OP_NONE else "")
        parts.append(s)
    return "|".join(parts)


def spec_name(sig: str) -> str:
    return "nmspec_" + hashlib.sha1(sig.encode()).hexdigest()[:16]


def _term_tables(p: L.NmPlan) -> str:
    """Compile-time term structure: per term cmp / norm / side ops, per atom variable + broadcast flags."""
    n = p.n_terms
    cmp_, norm, ops, var, tb, db = [], [], [], [], [], []
    for k in range(n):
        t = p.terms[k]
        cmp_.append(t.cmp); norm.append(t.norm); ops += [t.lhs.op, t.rhs.op]
        for side in (t.lhs, t.rhs):
            for live, a in ((True, side.a), (side.op != L.NM_OP_NONE, side.b)):
                if not live:
                    var.append(-2); tb.append(0); db.append(0)
                else:
                    var.append(a.var if a.var >= 0 else -1)
                    tb.append(int(a.var >= 0 and a.t_len == 1 and t.t_len > 1))
                    db.append(int(a.var >= 0 and a.d_len == 1 and t.d_len > 1))

    def tab(name, vals, args, index):
        vals = list(vals) or [0]
        return (f"  static constexpr int {name}({args}) {{ constexpr int V[] = {{{', '.join(map(str, vals))}}}; "
                f"return V[{index}]; }}\n")
    return (f"  static constexpr int NTERMS = {n};\n" + tab("t_cmp", cmp_, "int k", "k") + tab("t_norm", norm, "int k", "k")
            + tab("t_op", ops, "int k, int side", "2 * k + side") + tab("a_var", var, "int k, int w", "4 * k + w")
            + tab("a_tb", tb, "int k, int w", "4 * k + w") + tab("a_db", db, "int k, int w", "4 * k + w")
            + f"  static constexpr const char* term_signature() {{ return \"{term_signature_of(p)}\"; }}\n")


# ------------------------------------------------------------------------------- header generation
def _mlp_struct(name: str, m, present: bool) -> str:
    if not present:
        return (f"  struct {name} {{ static constexpr int NL = 0; static constexpr int sz(int) {{ return 1; }}\n"
                f"    static constexpr int ACT = 0, BOUNDS = 0; static constexpr bool BIAS = false; }};\n")
    sizes = ", ".join(str(m.sizes[i]) for i in range(m.n_layers + 1))
    return (f"  struct {name} {{ static constexpr int NL = {m.n_layers};\n"
            f"    static constexpr int sz(int i) {{ constexpr int S[] = {{{sizes}}}; return S[i]; }}\n"
            f"    static constexpr int ACT = {m.act}, BOUNDS = {m.bounds}; static constexpr bool BIAS = {'true' if m.has_bias else 'false'}; }};\n")


def _table(fn: str, vals: List[int]) -> str:
    vals = list(vals) or [0]
    return (f"  static constexpr int {fn}(int i) {{ constexpr int V[] = {{{', '.join(map(str, vals))}}}; "
            f"return V[i]; }}\n")


# Per-shape-class launch tuning (tile = trajectories per CTA, threads per CTA, __launch_bounds__ min
# CTAs/SM for the forward / adjoint kernel), found by measurement on B200 -- see DESIGN.md.
# NM_TUNE="tile=64,threads=128,minb_fwd=4,minb_bwd=4" overrides it for every class (experiments).
TUNING: Dict[str, Dict[str, int]] = {}
# substring rules: the small RKF neural-ODE test shape is forced onto the tensor-core forward kernel so that the
# parity suite covers it on a second shape class (its 16-wide layers would not pick it on throughput grounds)
# c4 (193-wide fan-in, 192 of them exogenous look-ahead features): the exogenous part of the first layer is hoisted out of the
# horizon loop into one GEMM before the rollout and one after the adjoint (nm_kernels.cuh, HoistedMlp) -- GPU parity
# green since round 2 (tests/test_hoist_gpu.py, then the whole -m gpu suite with the rule on)
TUNING_RULES = [("rhs4[3-16-16-2]", {"tc": 1}), ("pol[193-32-32-1]", {"hoist": 1})]


def parse_tune(text: str) -> Dict[str, int]:
    return {k.strip(): int(v) for k, v in (kv.split("=") for kv in text.split(",") if kv.strip())}


def tuning_for(sig: str) -> Dict[str, int]:
    env = os.environ.get("NM_TUNE")
    if env:
        return parse_tune(env)
    out = dict(TUNING.get(sig, {}))
    for pat, kv in TUNING_RULES:
        if pat in sig:
            out.update(kv)
    return out


def spec_header(p: L.NmPlan, tune_override: Dict[str, int] = None) -> str:
    sig = signature_of(p)
    name = spec_name(sig + "#" + term_signature_of(p))
    tune_map = tune_override if tune_override is not None else tuning_for(sig)
    tune = "".join(f"#define NM_TUNE_{k.upper()} {v}\n" for k, v in sorted(tune_map.items()))
    # BarrierLoss terms among the compiled ones?  (nm_terms.cuh compiles the barrier branch out of the unit otherwise)
    tune += f"#define NM_SPEC_BARRIER {int(any((p.terms[k].norm >> 4) & 15 for k in range(p.n_terms)))}\n"
    nd = p.exo[p.dist_exo].width if p.dist_exo >= 0 else 0
    segs = [p.segs[i] for i in range(p.n_segs)] if p.has_policy else []
    integ = p.integrator if p.dyn_kind == L.NM_DYN_ODE else 0
    out = [f"// generated by neuromancer_b200/build.py -- shape class\n// {sig}\n#pragma once\n", tune,
           f"#define NM_SPEC_NAME {name}\n", f"struct {name} {{\n",
           f"  static constexpr int NX = {p.nx}, NU = {p.nu}, NY = {p.ny}, ND = {nd};\n",
           _mlp_struct("Pol", p.policy, bool(p.has_policy)),
           _mlp_struct("RhsM", p.rhs_mlp, L.rhs_has_mlp(p.rhs_kind)),
           f"  static constexpr bool HAS_OUTPUT = {'true' if p.has_output else 'false'};\n",
           f"  static constexpr int NSEG = {len(segs)};\n",
           _table("seg_kind", [s.kind for s in segs]),
           _table("seg_index", [s.index for s in segs]),
           _table("seg_width", [s.width for s in segs]),
           _table("seg_sub", [(p.exo[s.index].width if s.kind in (L.NM_SEG_EXO, L.NM_SEG_PREVIEW) else s.width) for s in segs]),
           f"  static constexpr int DYN = {p.dyn_kind}, INTEG = {integ}, RHS = {p.rhs_kind}, DIST_EXO = {p.dist_exo}, ACT_EXO = {p.act_exo};\n",
           f"  static constexpr const char* signature() {{ return \"{sig}\"; }}\n", _term_tables(p), "};\n"]
    return "".join(out)


def write_spec(p: L.NmPlan) -> str:
    os.makedirs(SPEC_DIR, exist_ok=True)
    name = spec_name(signature_of(p) + "#" + term_signature_of(p))
    path = os.path.join(SPEC_DIR, name + ".h")
    text = spec_header(p)
    if not os.path.exists(path) or open(path).read() != text:
        with open(path, "w") as f:
            f.write(text)
    return path


# ------------------------------------------------------------------------------- compilation
def _sources_mtime() -> float:
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    files.append(os.path.join(INCLUDE, "nm_rollout.h"))
    return max(os.path.getmtime(f) for f in files)


def _run(cmd: List[str]) -> None:
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout[-8000:])
    if os.environ.get("NM_BUILD_VERBOSE"):
        print(res.stdout)


def compile_spec_object(header: str, force: bool = False, extra: List[str] = ()) -> str:
    os.makedirs(OBJ_DIR, exist_ok=True)
    name = os.path.splitext(os.path.basename(header))[0]
    obj = os.path.join(OBJ_DIR, name + ".o")
    newest = max(_sources_mtime(), os.path.getmtime(header))
    if force or not os.path.exists(obj) or os.path.getmtime(obj) < newest:
        _run([nvcc_path(), *NVCC_FLAGS, *extra, f'-DNM_SPEC_HEADER="{header}"', "-c",
              os.path.join(CSRC, "nm_spec_tu.cu"), "-o", obj])
    return obj


def build_variant(case_name: str, tune: str, out: str) -> str:
    """Experiment helper: a library holding ONE shape class built with the tuning string ``tune``
    (same syntax as NM_TUNE), written to ``out``.  bench.py picks it up through NM_LIB_PATH."""
    from . import configs
    os.makedirs(os.path.dirname(out), exist_ok=True)
    p = configs.build_case(case_name).plan().c
    text = spec_header(p, tune_override=parse_tune(tune))
    tag = hashlib.sha1((case_name + tune).encode()).hexdigest()[:10]
    vdir = os.path.join(OUT_DIR, "variants")
    os.makedirs(vdir, exist_ok=True)
    header = os.path.join(vdir, f"{spec_name(signature_of(p) + '#' + term_signature_of(p))}_{tag}.h")
    with open(header, "w") as f:
        f.write(text)
    obj = os.path.join(vdir, f"{tag}.o")
    _run([nvcc_path(), *NVCC_FLAGS, f'-DNM_SPEC_HEADER="{header}"', "-c", os.path.join(CSRC, "nm_spec_tu.cu"), "-o", obj])
    api_obj = os.path.join(OBJ_DIR, "nm_api.o")
    _run([nvcc_path(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out, api_obj, obj, "-lcudart", "-ldl"])
    return out


def build_library(plans: List[L.NmPlan], force: bool = False, verbose: bool = True) -> str:
    """Compile ``libnm_rollout.so`` holding the given shape classes."""
    os.makedirs(OBJ_DIR, exist_ok=True)
    headers = sorted({write_spec(p) for p in plans})
    api_obj = os.path.join(OBJ_DIR, "nm_api.o")
    jobs = []
    if force or not os.path.exists(api_obj) or os.path.getmtime(api_obj) < _sources_mtime():
        jobs.append(lambda: _run([nvcc_path(), *NVCC_FLAGS, "-c", os.path.join(CSRC, "nm_api.cu"), "-o", api_obj]))
    objs: Dict[str, str] = {}

    def spec_job(h):
        objs[h] = compile_spec_object(h, force=force)
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        futs = [ex.submit(j) for j in jobs] + [ex.submit(spec_job, h) for h in headers]
        for f in futs:
            f.result()
    all_objs = [api_obj] + [objs[h] for h in headers]
    lib = L.LIB_PATH
    if force or not os.path.exists(lib) or any(os.path.getmtime(o) > os.path.getmtime(lib) for o in all_objs):
        _run([nvcc_path(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib, *all_objs,
              "-lcudart", "-ldl"])
    if verbose:
        print(f"[neuromancer_b200.build] {lib}: {len(headers)} shape classes")
    return lib


def ensure_specialisation(plan) -> None:
    """Compile and load the shape class of ``plan`` if the library does not have it."""
    import ctypes as C
    lib = L.load()
    if lib.nm_has_kernel(C.byref(plan.c)):
        return
    header = write_spec(plan.c)
    name = os.path.splitext(os.path.basename(header))[0]
    so = os.path.join(OUT_DIR, f"nm_spec_{name}.so")
    if not os.path.exists(so) or os.path.getmtime(so) < _sources_mtime():
        obj = compile_spec_object(header)
        _run([nvcc_path(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", so, obj, "-lcudart"])
    L.check(lib.nm_load_plugin(so.encode()), "nm_load_plugin")


# ------------------------------------------------------------------------------- AOT registry
def aot_plans() -> List[L.NmPlan]:
    """Shape classes compiled ahead of time: the five benchmark configs + the test shapes."""
    from . import configs
    plans = []
    for name in configs.AOT_CASES:
        case = configs.build_case(name, batch=4, device="cpu")
        plans.append(case.plan().c)
    return plans


def main(argv=None) -> int:
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    args = ap.parse_args(argv)
    build_library(aot_plans(), force=args.force)
    return 0


if __name__ == "__main__":
    sys.exit(main())
