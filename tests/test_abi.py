"""C-ABI library: loads without a GPU, exports every symbol include/nm_rollout.h declares, structure
layouts match the ctypes mirror, shape-class signatures agree between C and Python.  No compute."""
import ctypes as C
import os
import re

import pytest

from neuromancer_b200 import _lib as L
from neuromancer_b200 import build, configs

HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "nm_rollout.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nm_[a-z_0-9]+)\s*\(", text)))


def test_header_functions_all_exported(nm_lib):
    names = declared_functions()
    assert "nm_rollout_forward" in names and "nm_rollout_backward" in names
    for fn in names:
        assert hasattr(nm_lib, fn), f"{fn} declared in nm_rollout.h but not exported"
    assert set(L.EXPORTED_SYMBOLS) <= set(names)


def test_abi_version_and_struct_sizes(nm_lib):
    assert nm_lib.nm_abi_version() == L.NM_ABI_VERSION
    sizes = (C.c_int * 7)()
    assert nm_lib.nm_struct_sizes(sizes, 7) == 7
    mirror = [C.sizeof(t) for t in (L.NmMlp, L.NmSegment, L.NmExo, L.NmAtom, L.NmSide, L.NmTerm, L.NmPlan)]
    assert list(sizes) == mirror


@pytest.mark.parametrize("name", configs.AOT_CASES)
def test_aot_shape_classes_registered(nm_lib, name):
    plan = configs.build_case(name).plan()
    assert L.signature(plan.c) == build.signature_of(plan.c)
    assert nm_lib.nm_has_kernel(C.byref(plan.c)) == 1
    info = nm_lib.nm_kernel_info(C.byref(plan.c)).decode()
    assert "tile=" in info and "threads=" in info
    f, b = C.c_size_t(), C.c_size_t()
    assert nm_lib.nm_rollout_workspace_bytes(C.byref(plan.c), 4096, C.byref(f), C.byref(b)) == L.NM_OK
    assert f.value > 0 and b.value >= 4 * plan.c.n_params


def test_bad_plans_are_rejected(nm_lib):
    plan = configs.build_case("c2").plan()
    p = L.NmPlan.from_buffer_copy(plan.c)
    p.abi_version = 999
    buf = C.create_string_buffer(512)
    assert nm_lib.nm_plan_signature(C.byref(p), buf, 512) == L.NM_ERR_ABI
    p = L.NmPlan.from_buffer_copy(plan.c)
    p.terms[0].lhs.a.t_len = 7          # does not broadcast against the term's 31 steps
    assert nm_lib.nm_plan_signature(C.byref(p), buf, 512) == L.NM_ERR_BAD_PLAN
    assert b"broadcast" in nm_lib.nm_last_error()
    p = L.NmPlan.from_buffer_copy(plan.c)
    p.policy.sizes[1] = 33              # a shape class nobody compiled
    assert nm_lib.nm_has_kernel(C.byref(p)) == 0
    f, b = C.c_size_t(), C.c_size_t()
    assert nm_lib.nm_rollout_workspace_bytes(C.byref(p), 16, C.byref(f), C.byref(b)) == L.NM_ERR_NO_KERNEL


def test_null_arguments_fail_without_touching_the_device(nm_lib):
    plan = configs.build_case("c1").plan()
    rc = nm_lib.nm_rollout_forward(C.byref(plan.c), None, None, None, None, None, None, None, None, None, 0, 8, None)
    assert rc == L.NM_ERR_BAD_PLAN


def test_neural_ode_classes_are_served_by_the_tensor_core_kernels(nm_lib):
    """The c3 shape class (60-wide hidden layers) is compiled onto the tcgen05 kernels, the small RKF test shape is
    forced onto them (build.TUNING_RULES) so that the parity suite covers a second class; narrow policies stay on
    the FFMA kernels.  Checkpoint buffer: stage derivatives k_0..k_{NS-2} of every step."""
    info = nm_lib.nm_kernel_info(C.byref(configs.build_case("c3").plan().c)).decode()
    assert "fwd[tcgen05" in info and "bwd[tcgen05" in info and "1xTF32-rn" in info
    info = nm_lib.nm_kernel_info(C.byref(configs.build_case("t_int_rkf_node").plan().c)).decode()
    assert "fwd[tcgen05" in info and "bwd[tcgen05" in info and "3xTF32" in info
    # narrow closed-loop policies (tiny fan-in / fan-out, hidden layers >= 16 wide) run in policy mode
    for name in ("c1", "c2", "t_vdp_rk4_tanh_clamp"):
        info = nm_lib.nm_kernel_info(C.byref(configs.build_case(name).plan().c)).decode()
        assert "fwd[tcgen05" in info and "bwd[tcgen05" in info and "1xTF32-rn" in info, (name, info)
    # c4 (193 inputs, 192 of them exogenous): first layer hoisted out of the loop, the rest runs in policy mode
    info = nm_lib.nm_kernel_info(C.byref(configs.build_case("c4").plan().c)).decode()
    assert "fwd[tcgen05" in info and "bwd[tcgen05" in info, info
    for name in ("c4p", "t_int_luther_vdp", "t_node_policy_euler"):
        assert "tcgen05" not in nm_lib.nm_kernel_info(C.byref(configs.build_case(name).plan().c)).decode(), name
    # the 128-wide policy of c5 runs on the wide-policy family (nm_wide_kernels.cuh): 3-product fp16 MMAs, weights
    # resident in shared memory, raw policy output as the checkpoint
    info = nm_lib.nm_kernel_info(C.byref(configs.build_case("c5").plan().c)).decode()
    assert "fwd[tcgen05 3xFP16 wide" in info and "bwd[tcgen05 3xFP16 wide" in info and "wgrad GEMM 1xFP16-rn" in info, info
    assert nm_lib.nm_rollout_checkpoint_bytes(C.byref(configs.build_case("c5").plan().c), 1000, C.byref(C.c_size_t())) == L.NM_OK
    plan = configs.build_case("c3").plan()
    nbytes = C.c_size_t()
    assert nm_lib.nm_rollout_checkpoint_bytes(C.byref(configs.build_case("c5").plan().c), 1000, C.byref(nbytes)) == L.NM_OK
    assert nbytes.value == 1000 * 200 * 6 * 4                                 # wide policy mode: o_t[6] per step
    assert nm_lib.nm_rollout_checkpoint_bytes(C.byref(plan.c), 1000, C.byref(nbytes)) == L.NM_OK
    assert nbytes.value == 1000 * plan.c.nsteps * 3 * plan.c.nx * 4          # RK4: 3 kept stages
    # policy mode keeps the raw policy output o_t of every step
    assert nm_lib.nm_rollout_checkpoint_bytes(C.byref(configs.build_case("c2").plan().c), 1000, C.byref(nbytes)) == L.NM_OK
    assert nbytes.value == 1000 * 30 * 2 * 4
    # hoisted c4: the first-layer pre-activation E[32] and the raw policy output o[1] of every step
    assert nm_lib.nm_rollout_checkpoint_bytes(C.byref(configs.build_case("c4").plan().c), 1000, C.byref(nbytes)) == L.NM_OK
    assert nbytes.value == 1000 * 96 * (32 + 1) * 4
    assert nm_lib.nm_rollout_checkpoint_bytes(C.byref(configs.build_case("c4p").plan().c), 1000, C.byref(nbytes)) == L.NM_OK
    assert nbytes.value == 0


def test_constraint_matrix_widths_and_exchange_argument_checks(nm_lib):
    """No GPU needed: nm_constraint_matrices answers the width query without a launch (B = 0); nm_grad_allreduce validates
    its arguments before touching the device."""
    plan = configs.build_case("c2").plan()
    w = (C.c_int64 * 3)()
    assert nm_lib.nm_constraint_matrices(C.byref(plan.c), None, None, None, None, None, None, None, None, None, None, w, 0, None) == L.NM_OK
    cons = [plan.c.terms[k] for k in range(plan.c.n_terms) if not plan.c.terms[k].is_objective]
    total = sum(t.t_len * t.d_len for t in cons)
    eq = sum(t.t_len * t.d_len for t in cons if t.cmp == L.NM_CMP_EQ)
    assert (w[0], w[1], w[2]) == (total, eq, total - eq) and total > 0
    assert nm_lib.nm_grad_allreduce_buffer_floats(1290) == 2 * 1344 + 64            # halves padded to 64 floats + flag words
    peers = (C.c_void_p * 2)(0, 0)
    assert nm_lib.nm_grad_allreduce(peers, None, None, 10, 0, 2, 1, 10, 1.0, None) == L.NM_ERR_BAD_PLAN      # null output
    out = C.c_void_p(4096)
    assert nm_lib.nm_grad_allreduce(peers, None, out, 10, 0, 2, 0, 10, 1.0, None) == L.NM_ERR_BAD_PLAN       # epoch 0
    assert nm_lib.nm_grad_allreduce(peers, None, out, 10, 2, 2, 1, 10, 1.0, None) == L.NM_ERR_BAD_PLAN       # rank out of range
    assert nm_lib.nm_grad_allreduce(peers, None, out, 10, 0, 2, 1, 10, 1.0, None) == L.NM_ERR_BAD_PLAN       # null peer buffers
    assert b"nm_grad_allreduce" in nm_lib.nm_last_error()
