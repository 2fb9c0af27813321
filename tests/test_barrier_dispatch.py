"""A plan with BarrierLoss terms runs only on a specialisation compiled with its own term structure (the barrier code lives in
the compiled term evaluators of such a class and in the generic kernels of nm_api.cu, nowhere else): the C ABI refuses to hand
it to another class of the same shape, and the Python side builds the class on first use."""
import ctypes as C

from neuromancer_b200 import BarrierLoss, Problem, _lib as L, configs, variable


def test_barrier_plan_needs_its_compiled_term_class():
    lib = L.load()
    case = configs.build_case("t_barrier_softexp")                     # compiled ahead of time with its terms
    assert lib.nm_has_kernel(C.byref(case.plan().c)) == 1
    info = lib.nm_kernel_info(C.byref(case.plan().c)).decode()
    assert "fused_terms=1" in info
    # the same shape with ANOTHER barrier term structure: the shape class exists, this term class does not
    x, u = variable("x"), variable("u")
    lo = 1.5 * (x > -2.0)
    umin = (u > -0.9) ^ 2
    lo.name, umin.name = "lo2", "umin2"
    other = Problem([case.system], BarrierLoss([], [lo, umin], barrier="log", upper_bound=3.0))
    batch = case.make_batch(4, "cpu", 0)
    plan = case.system.plan_for(case.system.init(dict(batch)), loss=other.loss)
    assert plan.barrier and plan.c.n_terms == 2
    f, b = C.c_size_t(), C.c_size_t()
    assert lib.nm_has_kernel(C.byref(plan.c)) == 0
    assert lib.nm_rollout_workspace_bytes(C.byref(plan.c), 64, C.byref(f), C.byref(b)) == L.NM_ERR_NO_KERNEL
    assert b"BarrierLoss plan" in lib.nm_last_error()
    # without the barrier bits the very same terms are served by the shape class through the runtime descriptors
    for k in range(plan.c.n_terms):
        plan.c.terms[k].norm &= 15
    assert lib.nm_has_kernel(C.byref(plan.c)) == 1
    assert lib.nm_rollout_workspace_bytes(C.byref(plan.c), 64, C.byref(f), C.byref(b)) == L.NM_OK
