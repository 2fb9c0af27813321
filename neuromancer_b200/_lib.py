"""
ctypes binding of ``libnm_rollout.so`` -- the C ABI declared in ``include/nm_rollout.h``.

The structures below mirror the header field by field (``tests/test_abi.py`` checks ``sizeof``
against the values the library reports).  There is no fallback: if the shared object has not been
built (``python -m neuromancer_b200.build``) loading raises, and every rollout fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

NM_ABI_VERSION = 7
NM_MAX_LAYERS, NM_MAX_SEGS, NM_MAX_EXO, NM_MAX_TERMS, NM_MAX_DIM = 8, 8, 8, 16, 16

NM_OK, NM_ERR_BAD_PLAN, NM_ERR_NO_KERNEL, NM_ERR_CUDA, NM_ERR_WORKSPACE, NM_ERR_ABI = 0, -1, -2, -3, -4, -5

NM_BOUNDS_NONE, NM_BOUNDS_SIGMOID_SCALE, NM_BOUNDS_RELU_CLAMP = 0, 1, 2
NM_SEG_STATE, NM_SEG_OUTPUT, NM_SEG_EXO, NM_SEG_PREVIEW = 0, 1, 2, 3
NM_PAD_CONSTANT, NM_PAD_REPLICATE, NM_PAD_CIRCULAR, NM_PAD_REFLECT = 0, 1, 2, 3
NM_DYN_SSM, NM_DYN_ODE = 0, 1
NM_INT_EULER, NM_INT_RK4, NM_INT_RK2, NM_INT_EULER_TRAP, NM_INT_RK4_TRAP, NM_INT_LUTHER, NM_INT_RKF = 0, 1, 2, 3, 4, 5, 6
NM_INT_LEAPFROG, NM_INT_YOSHIDA4 = 7, 8
NM_RHS_NONE, NM_RHS_TWOTANK, NM_RHS_VDP, NM_RHS_CVDP, NM_RHS_MLP, NM_RHS_LINEAR = 0, 1, 2, 3, 4, 5
NM_RHS_BRUSS_HYBRID, NM_RHS_LV_HYBRID = 12, 13


def rhs_has_mlp(kind: int) -> bool:
    """NM_RHS_HAS_MLP of include/nm_rollout.h: a neural or a hybrid (grey-box) right-hand side"""
    return kind in (NM_RHS_MLP, NM_RHS_BRUSS_HYBRID, NM_RHS_LV_HYBRID)
NM_VAR_CONST, NM_VAR_X, NM_VAR_U, NM_VAR_Y, NM_VAR_EXO0 = -1, 0, 1, 2, 3
NM_OP_NONE, NM_OP_ADD, NM_OP_SUB, NM_OP_MUL = 0, 1, 2, 3
NM_CMP_EQ, NM_CMP_LT, NM_CMP_GT = 0, 1, 2
NM_BARRIERS = {"log10": 1, "log": 2, "inverse": 3, "softexp": 4, "softlog": 5, "expshift": 6}


class NmMlp(C.Structure):
    _fields_ = [("n_layers", C.c_int32), ("sizes", C.c_int32 * (NM_MAX_LAYERS + 1)),
                ("act", C.c_int32), ("act_param", C.c_float), ("has_bias", C.c_int32),
                ("bounds", C.c_int32), ("bmin", C.c_float * NM_MAX_DIM), ("bmax", C.c_float * NM_MAX_DIM),
                ("param_offset", C.c_int64), ("trainable", C.c_int32), ("_pad", C.c_int32)]


class NmSegment(C.Structure):
    _fields_ = [("kind", C.c_int32), ("index", C.c_int32), ("width", C.c_int32), ("preview_len", C.c_int32)]


class NmExo(C.Structure):
    _fields_ = [("width", C.c_int32), ("steps", C.c_int32)]


class NmAtom(C.Structure):
    _fields_ = [("var", C.c_int32), ("t_start", C.c_int32), ("t_len", C.c_int32),
                ("d_start", C.c_int32), ("d_len", C.c_int32), ("scale", C.c_float),
                ("offset", C.c_float), ("_pad", C.c_int32)]


class NmSide(C.Structure):
    _fields_ = [("op", C.c_int32), ("_pad", C.c_int32), ("a", NmAtom), ("b", NmAtom)]


class NmTerm(C.Structure):
    _fields_ = [("cmp", C.c_int32), ("norm", C.c_int32), ("weight", C.c_float),
                ("is_objective", C.c_int32), ("t_len", C.c_int32), ("d_len", C.c_int32),
                ("lhs", NmSide), ("rhs", NmSide)]


_MAT = C.c_float * (NM_MAX_DIM * NM_MAX_DIM)


class NmPlan(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("nx", C.c_int32), ("nu", C.c_int32), ("ny", C.c_int32),
                ("nsteps", C.c_int32), ("has_policy", C.c_int32), ("has_output", C.c_int32),
                ("n_segs", C.c_int32), ("segs", NmSegment * NM_MAX_SEGS), ("policy", NmMlp),
                ("dyn_kind", C.c_int32), ("integrator", C.c_int32), ("rhs_kind", C.c_int32),
                ("dist_exo", C.c_int32), ("act_exo", C.c_int32), ("h", C.c_float), ("rhs_const", C.c_float * 8),
                ("rhs_mlp", NmMlp), ("A", _MAT), ("Bm", _MAT), ("E", _MAT), ("C", _MAT),
                ("n_exo", C.c_int32), ("exo", NmExo * NM_MAX_EXO), ("n_terms", C.c_int32),
                ("terms", NmTerm * NM_MAX_TERMS), ("n_params", C.c_int64),
                ("pad_mode", C.c_int32), ("pad_value", C.c_float),
                ("barrier_alpha", C.c_float), ("barrier_shift", C.c_float), ("barrier_ub", C.c_float), ("reserved0", C.c_int32)]


_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.path.join(_HERE, "_C")
LIB_PATH = os.environ.get("NM_LIB_PATH") or os.path.join(LIB_DIR, "libnm_rollout.so")

_lib = None


class NmError(RuntimeError):
    pass


def _declare(lib):
    P = C.POINTER
    fp, vp = C.c_void_p, C.c_void_p
    lib.nm_abi_version.restype = C.c_int
    lib.nm_last_error.restype = C.c_char_p
    lib.nm_plan_signature.argtypes = [P(NmPlan), C.c_char_p, C.c_int]
    lib.nm_plan_signature.restype = C.c_int
    lib.nm_plan_term_signature.argtypes = [P(NmPlan), C.c_char_p, C.c_int]
    lib.nm_plan_term_signature.restype = C.c_int
    lib.nm_has_kernel.argtypes = [P(NmPlan)]
    lib.nm_has_kernel.restype = C.c_int
    lib.nm_num_kernels.restype = C.c_int
    lib.nm_kernel_signature.argtypes = [C.c_int]
    lib.nm_kernel_signature.restype = C.c_char_p
    lib.nm_load_plugin.argtypes = [C.c_char_p]
    lib.nm_load_plugin.restype = C.c_int
    lib.nm_rollout_workspace_bytes.argtypes = [P(NmPlan), C.c_int64, P(C.c_size_t), P(C.c_size_t)]
    lib.nm_rollout_workspace_bytes.restype = C.c_int
    lib.nm_rollout_checkpoint_bytes.argtypes = [P(NmPlan), C.c_int64, P(C.c_size_t)]
    lib.nm_rollout_checkpoint_bytes.restype = C.c_int
    lib.nm_rollout_forward.argtypes = [P(NmPlan), fp, fp, P(C.c_void_p), fp, fp, fp, fp, fp, vp,
                                       C.c_size_t, C.c_int64, vp]
    lib.nm_rollout_forward.restype = C.c_int
    lib.nm_rollout_backward.argtypes = [P(NmPlan), fp, P(C.c_void_p), fp, fp, fp, fp, fp, fp, fp, fp,
                                        fp, fp, vp, C.c_size_t, C.c_int64, C.c_int64, vp]
    lib.nm_rollout_backward.restype = C.c_int
    lib.nm_last_launch_count.restype = C.c_int
    lib.nm_struct_sizes.argtypes = [P(C.c_int), C.c_int]
    lib.nm_struct_sizes.restype = C.c_int
    lib.nm_kernel_info.argtypes = [P(NmPlan)]
    lib.nm_kernel_info.restype = C.c_char_p
    lib.nm_adamw_workspace_bytes.argtypes = [C.c_int64]
    lib.nm_adamw_workspace_bytes.restype = C.c_size_t
    lib.nm_adamw_step.argtypes = [fp, fp, fp, fp, vp, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                  C.c_int64, C.c_float, fp, vp, C.c_size_t, vp]
    lib.nm_adamw_step.restype = C.c_int
    lib.nm_constraint_matrices.argtypes = [P(NmPlan), P(C.c_void_p), fp, fp, fp, fp, fp, fp, fp, fp, fp, P(C.c_int64), C.c_int64, vp]
    lib.nm_constraint_matrices.restype = C.c_int
    lib.nm_grad_allreduce_buffer_floats.argtypes = [C.c_int64]
    lib.nm_grad_allreduce_buffer_floats.restype = C.c_size_t
    lib.nm_grad_allreduce.argtypes = [P(C.c_void_p), vp, vp, C.c_int64, C.c_int, C.c_int, C.c_uint32, C.c_int64, C.c_float, vp]
    lib.nm_grad_allreduce.restype = C.c_int


EXPORTED_SYMBOLS = [
    "nm_abi_version", "nm_last_error", "nm_plan_signature", "nm_has_kernel", "nm_num_kernels",
    "nm_kernel_signature", "nm_load_plugin", "nm_rollout_workspace_bytes", "nm_rollout_forward",
    "nm_rollout_backward", "nm_last_launch_count", "nm_kernel_info", "nm_struct_sizes",
    "nm_plan_term_signature", "nm_rollout_checkpoint_bytes", "nm_constraint_matrices",
    "nm_grad_allreduce", "nm_grad_allreduce_buffer_floats",
]


def load():
    """Load (once) and return the CDLL.  Raises ``NmError`` when the extension is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NmError(
            f"{LIB_PATH} not found: the sm_100a extension has not been built. Run "
            "`python -m neuromancer_b200.build` (or __graft_entry__.build()). neuromancer_b200 has "
            "no CPU or eager fallback for rollouts.")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    _declare(lib)
    got = lib.nm_abi_version()
    if got != NM_ABI_VERSION:
        raise NmError(f"libnm_rollout.so ABI {got} != python binding ABI {NM_ABI_VERSION}; rebuild")
    _lib = lib
    # specialisations built on demand for shapes outside the ahead-of-time registry
    # (a plugin older than the library was built from older kernel sources -- its checkpoint / workspace layouts may differ:
    # it is left alone and rebuilt on demand by build.ensure_specialisation)
    if os.path.isdir(LIB_DIR) and not os.environ.get("NM_LIB_PATH"):
        built = os.path.getmtime(LIB_PATH)
        for fn in sorted(os.listdir(LIB_DIR)):
            path = os.path.join(LIB_DIR, fn)
            if fn.startswith("nm_spec_") and fn.endswith(".so") and os.path.getmtime(path) >= built:
                lib.nm_load_plugin(path.encode())
    return lib


def check(status, what=""):
    if status != NM_OK:
        msg = load().nm_last_error().decode(errors="replace")
        raise NmError(f"{what} failed with status {status}: {msg}")


def term_signature(plan: NmPlan) -> str:
    buf = C.create_string_buffer(4096)
    n = load().nm_plan_term_signature(C.byref(plan), buf, len(buf))
    if n < 0:
        check(n, "nm_plan_term_signature")
    return buf.value.decode()


def signature(plan: NmPlan) -> str:
    buf = C.create_string_buffer(1024)
    n = load().nm_plan_signature(C.byref(plan), buf, len(buf))
    if n < 0:
        check(n, "nm_plan_signature")
    return buf.value.decode()
