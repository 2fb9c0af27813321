"""The FFMA kernel templates (neuromancer_b200/csrc/nm_kernels.cuh) compiled for the HOST and run against the oracle.

tests/emu/cuda_emu.h emulates the CUDA execution model with one OS thread per CUDA thread of one block at a time
(std::barrier for __syncthreads); tests/emu/ffma_emu.cpp launches weight preparation, rollout and adjoint kernels the
way nm_spec_tu.cu does and writes trajectories, term sums and the parameter gradient.  What this checks without a GPU
is the kernels' LOGIC -- index math, staging layouts, barrier placement (a missing barrier is a data race between the
OS threads too), the persistent tile loop, the per-CTA gradient blocks -- not their performance and not the
tensor-core kernels (inline PTX).  It is how the flag-guarded first-layer hoist (HoistedMlp, nm_exo_gemm_*_kernel) was
validated at the end of round 1, after the GPU budget was spent."""
import os
import shutil
import struct
import subprocess
import sys

import numpy as np
import pytest
import torch

import helpers as H
from neuromancer_b200 import build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ (C++20)")


def compile_emulator(name, out, defines=(), tsan=False, src="ffma_emu.cpp"):
    hdr = build.write_spec(H.configs.build_case(name).plan().c)
    cmd = ["g++", "-std=c++20", "-O1", "-w", "-fpermissive", "-DNM_HOST_EMU", *(["-g", "-fsanitize=thread"] if tsan else []), "-I", os.path.join(ROOT, "tests", "emu"),
           "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "neuromancer_b200", "csrc"),
           "-I", "/usr/local/cuda/include", f'-DNM_SPEC_HEADER="{hdr}"', *[f"-D{d}" for d in defines],
           os.path.join(ROOT, "tests", "emu", src), "-o", out, "-lpthread"]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if tsan and res.returncode != 0 and "tsan" in res.stdout.lower():
        pytest.skip("libtsan not available")
    assert res.returncode == 0, res.stdout[-4000:]


def run_emulated(name, B, exe, tmp):
    case = H.configs.build_case(name, seed=3)
    batch = case.make_batch(B, "cpu", 7)
    x_in = case.system.init(dict(batch))
    plan = case.system.plan_for(x_in, loss=case.problem.loss)
    p = plan.c
    flat = np.zeros(int(p.n_params), np.float32)
    value = lambda q: (q.materialise() if hasattr(q, "materialise") else q).detach()    # derived entries: evaluated now
    for q, (off, n) in zip(plan.params, plan.param_slices):
        flat[off:off + n] = value(q).numpy().reshape(-1)
    lay = case.system.layout()
    fin, fout = os.path.join(tmp, "in.bin"), os.path.join(tmp, "out.bin")
    with open(fin, "wb") as f:
        f.write(struct.pack("q", B))
        f.write(bytes(p))
        f.write(flat.tobytes())
        f.write(x_in[lay.state_key][:, 0, :].contiguous().numpy().astype(np.float32).tobytes())
        for k in plan.exo_keys:
            f.write(x_in[k].contiguous().numpy().astype(np.float32).tobytes())
    res = subprocess.run([exe, fin, fout], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert "ThreadSanitizer" not in res.stdout, res.stdout[:3000]      # (only the -fsanitize=thread builds can say so)
    assert res.returncode == 0, res.stdout[-2000:]
    raw = open(fout, "rb").read()
    T, nx, nu, ny, pos = p.nsteps, p.nx, p.nu, p.ny, 0

    def take(n, dt=np.float32):
        nonlocal pos
        a = np.frombuffer(raw, dt, n, pos).copy()
        pos += n * np.dtype(dt).itemsize
        return a
    traj = {lay.state_key: take(B * (T + 1) * nx).reshape(B, T + 1, nx), lay.action_key: take(B * T * nu).reshape(B, T, nu),
            lay.output_key: take(B * T * ny).reshape(B, T, ny)}
    tsum, grad = take(p.n_terms, np.float64), take(int(p.n_params))
    # the oracle differentiates by the module parameters; a derived entry (white-box constants of a system-ID plant)
    # stands for its module's parameters, the kernel's gradient w.r.t. it is chained through its expression
    leaves = []
    for q in plan.params:
        leaves += q.own_parameters() if hasattr(q, "materialise") else [q]
    ref = H.oracle_run(name, [q.detach().clone() for q in leaves], batch)
    for k, v in ref["traj"].items():
        assert H.rel_err(torch.from_numpy(traj[k]), v) <= 1e-5, (name, B, k)
    loss = sum(p.terms[k].weight * tsum[k] / (B * p.terms[k].t_len * p.terms[k].d_len) for k in range(p.n_terms))
    assert abs(loss - float(ref["loss"].detach())) <= 1e-4 * abs(float(ref["loss"].detach())), (name, B, loss)
    trainable = [q for q in leaves if q.requires_grad]
    want = lambda q: ref["grads"][[id(z) for z in trainable].index(id(q))]
    for q, (off, n) in zip(plan.params, plan.param_slices):
        g = torch.from_numpy(grad[off:off + n]).view(q.shape)
        if hasattr(q, "materialise"):
            mine = [z for z in q.own_parameters() if z.requires_grad]
            for z, gz in zip(mine, torch.autograd.grad(q.materialise(), mine, grad_outputs=g, allow_unused=True)):
                gz = torch.zeros_like(z) if gz is None else gz
                assert H.rel_err(gz, want(z)) <= 1e-4, (name, B, "constants", H.rel_err(gz, want(z)))
        elif q.requires_grad:
            assert H.rel_err(g, want(q)) <= 1e-4, (name, B, tuple(q.shape), H.rel_err(g, want(q)))


# (case, defines): tile / thread shapes as nm_spec_tu.cu picks them, weights read from the prepared global block
PLAIN = [("c4", ("EMU_TILE=128", "EMU_THREADS=256"), (1, 150)),      # y = Cx node, disturbance, LeakyReLU, sigmoid bounds, 193-wide fan-in
         ("c2", ("EMU_TILE=128", "EMU_THREADS=128"), (1, 150)),      # exact FFMA fallback of the tensor-core class: GELU, TwoTank RK4, 5 terms
         ("c1", ("EMU_TILE=128", "EMU_THREADS=128"), (150,)),        # linear SSM, ReLU masks, 4 hidden layers
         ("c3", ("EMU_TILE=64", "EMU_THREADS=256"), (70,)),          # neural ODE: MLP right-hand side inside RK4, adjoint recomputing its stages
         ("c5", ("EMU_TILE=32", "EMU_THREADS=256"), (40,)),          # 128-wide layers, weight gradients spilled to the CTA's gradient block, coupled VdP ring
         ("t_ssm_softplus", ("EMU_TILE=64", "EMU_THREADS=128"), (1, 150)),     # state after the exogenous segment, softplus derivative buffers
         ("t_preview_reflect", ("EMU_TILE=128", "EMU_THREADS=128"), (150,)),   # SystemPreview window gather with reflect padding
         ("t_node_policy_euler", ("EMU_TILE=64", "EMU_THREADS=128"), (100,)),  # MLP policy AND MLP right-hand side
         ("t_int_leapfrog_node", ("EMU_TILE=64", "EMU_THREADS=128"), (1, 100)),  # second-order scheme: positions / velocities, a0 enters twice
         ("t_int_yoshida_node", ("EMU_TILE=64", "EMU_THREADS=128"), (100,)),     # 4th-order Yoshida composition, three evaluations
         ("t_sysid_twotank", ("EMU_TILE=128", "EMU_THREADS=128"), (1, 150)),     # gradient w.r.t. the white-box constants c1, c2 (clip / sqrt conventions)
         ("t_sysid_cstr", ("EMU_TILE=128", "EMU_THREADS=128"), (150,)),          # ten parameters behind seven derived coefficient groups
         ("t_sysid_duffing", ("EMU_TILE=64", "EMU_THREADS=64"), (100,)),         # cos(omega t): gradient w.r.t. a frequency
         ("t_sysid_vdp_policy", ("EMU_TILE=128", "EMU_THREADS=128"), (150,)),    # policy weights AND the plant's mu trained together
         ("t_ude_brusselator", ("EMU_TILE=64", "EMU_THREADS=128"), (1, 100)),    # grey-box: MLP 2-20-20-1 inside the Brusselator, weights + alpha, beta
         ("t_ude_lotka", ("EMU_TILE=64", "EMU_THREADS=128"), (100,)),            # grey-box Lotka-Volterra, four constants next to the MLP
         ("t_barrier_log10", ("EMU_TILE=128", "EMU_THREADS=128"), (150,)),       # BarrierLoss terms: barrier value / derivative on satisfied entries
         ("t_barrier_softexp", ("EMU_TILE=128", "EMU_THREADS=128"), (150,)),
         ("t_barrier_expshift", ("EMU_TILE=128", "EMU_THREADS=128"), (150,))]


@pytest.mark.parametrize("name,defines,batches", PLAIN, ids=[c[0] for c in PLAIN])
def test_ffma_kernels_on_the_host_match_the_oracle(name, defines, batches, tmp_path):
    exe = str(tmp_path / "emu")
    compile_emulator(name, exe, defines)
    for B in batches:                                                  # ragged single tile / several tiles on two CTAs
        run_emulated(name, B, exe, str(tmp_path))


@pytest.mark.parametrize("alias", [False, True])
def test_first_layer_hoist_on_the_host_matches_the_oracle(alias, tmp_path):
    """c4 with the exogenous part of its first layer hoisted into the two GEMM kernels (NM_TUNE_HOIST=1); `alias`:
    dZ_1 overwrites E in place, the no-checkpoint configuration of nm_spec_tu.cu."""
    exe = str(tmp_path / "emu_hoist")
    compile_emulator("c4", exe, ("EMU_TILE=128", "EMU_THREADS=256", "NM_TUNE_HOIST=1") + (("EMU_ALIAS_DZ1",) if alias else ()))
    for B in (1, 70, 300):
        run_emulated("c4", B, exe, str(tmp_path))


RACE = [("c2", ("EMU_TILE=128", "EMU_THREADS=128"), 130), ("c1", ("EMU_TILE=128", "EMU_THREADS=128"), 130),
        ("c4", ("EMU_TILE=128", "EMU_THREADS=256", "NM_TUNE_HOIST=1", "EMU_ALIAS_DZ1"), 130)]


@pytest.mark.parametrize("name,defines,B", RACE, ids=["c2", "c1", "c4-hoist"])
def test_ffma_kernels_have_no_data_races_under_thread_sanitizer(name, defines, B, tmp_path):
    """The same host build under -fsanitize=thread: an unsynchronised shared-memory (or global) hand-over between the
    threads of a block -- a missing __syncthreads() -- is a data race between the emulating OS threads and is reported
    (checked with a deliberately broken kernel when this was written).  All ten classes of the list above were swept
    once; three stay in the suite."""
    exe = str(tmp_path / "emu_tsan")
    compile_emulator(name, exe, defines, tsan=True)
    run_emulated(name, B, exe, str(tmp_path))


# ------------------------------------------------------------------------------------------------
# tensor-core kernels (nm_tc_kernels.cuh) on the host: tests/emu/tc_ptx_emu.h models tensor memory (128 lanes x 512
# columns), tcgen05.mma on tf32-truncated operands read through the real shared-memory descriptors (interleaved K-major
# and SWIZZLE_128B), tcgen05.ld/st, mbarriers with transaction counts, bulk TMA copies and elect.sync; the warps of a
# block are OS threads, __syncwarp is a real 32-thread barrier.  Owner / control-warp protocol, operand and
# transposed-tile layouts, ring-slot hand-over, weight-gradient flushes -- all of it runs, forward and adjoint.
TC = [("c2", 130),                      # policy mode, SOLO owners (two tiles on two CTAs), GELU derivative stash, single-pass TF32 weight gradients
      ("c1", 130),                      # policy mode, three tensor-core layers through the two-slot weight ring, ReLU masks
      ("t_vdp_rk4_tanh_clamp", 130),    # policy mode, tanh, relu-clamp bounds
      ("t_int_rkf_node", 130),          # right-hand-side mode, six stages per step, 3xTF32 weight gradients
      ("t_ssm_softplus", 130),          # policy mode with an output node y = C x, a disturbance and [y | exo | x] inputs (FFMA by default: 10-wide)
      ("c3", 60)]                       # right-hand-side mode, PAIR owners (named barriers), 60-wide layers


@pytest.mark.parametrize("name,B", TC, ids=[c[0] for c in TC])
def test_tensor_core_kernels_on_the_host_match_the_oracle(name, B, tmp_path):
    exe = str(tmp_path / "emu_tc")
    compile_emulator(name, exe, src="tc_emu.cpp")
    run_emulated(name, B, exe, str(tmp_path))


def test_tensor_core_kernels_have_no_data_races_under_thread_sanitizer(tmp_path):
    """Owner <-> control-warp hand-overs of the tcgen05 kernels (operands in tensor memory, transposed tiles and the
    weight ring in shared memory) under -fsanitize=thread.  The emulated MMA runs when it is issued, so what this sees
    are read-after-write hazards (an MMA issued before the owners' stores are ordered by the mbarrier); write-after-read
    hazards against a still-running asynchronous MMA are outside this model."""
    exe = str(tmp_path / "emu_tc_tsan")
    compile_emulator("c2", exe, ("EMU_DEADLOCK_SECONDS=120",), tsan=True, src="tc_emu.cpp")
    run_emulated("c2", 130, exe, str(tmp_path))


def test_first_layer_hoist_on_the_tensor_core_kernels_matches_the_oracle(tmp_path):
    """c4 with NM_TUNE_HOIST=1 is a policy-mode tensor-core class: E from nm_exo_gemm_fwd_kernel at the head of the
    checkpoint buffer, a 1-input 32-wide policy with E as per-step pre-activation, y = C x and the disturbance as register
    math, dZ_1 stored by the adjoint's layer-1 epilogue, dW_0[:, exo] from nm_exo_gemm_bwd_kernel."""
    exe = str(tmp_path / "emu_tc_hoist")
    compile_emulator("c4", exe, ("NM_TUNE_HOIST=1",), src="tc_emu.cpp")
    for B in (1, 130):
        run_emulated("c4", B, exe, str(tmp_path))
    # ... and with both hoist GEMMs on the tensor cores: E from nm_exo_gemm_tc_fwd_kernel (row slices as K-major operand
    # + remainder tiles in shared memory, 3-product MMAs, bias in the epilogue), dW_0[:, exo] from
    # nm_exo_gemm_tc_bwd_kernel (transposed tf32 tiles in the 128-byte swizzle, contraction over rows, accumulators
    # flushed every 20 tiles; single-pass TF32 like the adjoint's weight gradients)
    exe2 = str(tmp_path / "emu_tc_hoist_tcgemm")
    compile_emulator("c4", exe2, ("NM_TUNE_HOIST=1", "EMU_TC_GEMM"), src="tc_emu.cpp")
    for B in (1, 130):
        run_emulated("c4", B, exe2, str(tmp_path))
