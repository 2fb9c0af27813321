#!/usr/bin/env python
"""
bench.py -- DPC train-step throughput of the fused rollout (forward + adjoint [+ gradient all-reduce]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c5] [--batch B] [--impl reference]

Contract (see the task statement): rank 0 prints ONE JSON line.  A "step" is one pass of the hot path over one batch of
synthetic input: nm_rollout_forward + nm_rollout_backward on device-resident buffers (`value`), timed with CUDA events,
L2 flushed between timed steps, max over ranks.  The default workload is the headline configuration of BASELINE.json:
**C5** (24-128x4-6 GELU policy, 12-state coupled Van der Pol + RK4, nsteps 200) at a GLOBAL batch of 4,194,304
trajectories -- strong scaling over 1/2/4/8 GPUs.  `e2e` is the same metric through the public API
(`Problem.forward` + `.backward()`) with every batch copied from pinned HOST memory inside the timed region and the
loss read back every step.  `--impl reference` times the reference's own CPU implementation of the path on the host
cores: the UNMODIFIED reference modules (oracle/_ref, see oracle/make_ref.py) when they are present, else the oracle
port (pinned bit-exactly to the reference).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "DPC train-step trajectory-steps/sec (fwd+bwd)"
UNIT = "trajectory-steps/s"
# CPU arm: batch per workload (the largest the eager reference holds an autograd tape for)
CPU_REF_BATCH = {"c1": 3333, "c2": 65536, "c3": 8192, "c4": 16384, "c4p": 16384, "c5": 4096}
CPU_PORT_BATCH = {"c1": 3333, "c2": 8192, "c3": 2048, "c4": 2048, "c4p": 2048, "c5": 256}
E2E_MAX_BATCH = 1048576          # per GPU: the public-API leg double-buffers its host batches (c5: 10 GB of `r` each)
DESCRIPTION = {"c1": "DPC double integrator, MLP 2-20-20-20-20-1 ReLU, linear SSM",
               "c2": "DPC reference tracking, TwoTank ODE + RK4, MLP_bounds 4-32-32-2 GELU",
               "c3": "neural-ODE system ID, RK4 over MLP RHS 2-60-60-60-2 ReLU",
               "c4": "building thermal DPC, linear SSM nx=4, 48-step preview, MLP_bounds 193-32-32-1",
               "c4p": "c4 with SystemPreview windows over the raw signals (in-kernel gather)",
               "c5": "parametric nonlinear DPC, 12-state coupled VdP + RK4, MLP_bounds 24-128x4-6 GELU, box + terminal constraints"}


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", os.environ.get("NM_CLOCK_MS", "50"), "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------- CPU arm
def _reference_src():
    """Unmodified reference modules: the build container's /root/reference, else the copy oracle/make_ref.py leaves
    under oracle/_ref (git-ignored, travels to the GPU box with the snapshot)."""
    for cand in (os.environ.get("NM_REFERENCE_SRC"), os.path.join(ROOT, "oracle", "_ref", "src"), "/root/reference/src"):
        if cand and os.path.isdir(os.path.join(cand, "neuromancer")):
            return cand
    return None


def time_cpu_reference(name, steps, warmup, nsteps, budget_s=45.0):
    """fwd+bwd of the reference's CPU path on all host cores, on a bounded sample of the workload."""
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    src = _reference_src()
    kind = "reference" if src is not None else "port"
    if kind == "reference":
        os.environ["NM_REFERENCE_SRC"] = src
        from oracle import pin_oracle, ref_shim
        ref_shim.REFERENCE_SRC = src
        ns = ref_shim.load()
        torch.manual_seed(0)
        builder = pin_oracle.REF_BUILDERS[name][0]
        problem, params, data_fn, _, _ = builder(ns)
        B = CPU_REF_BATCH.get(name, 1024)
        g = torch.Generator(device="cpu"); g.manual_seed(4321)
        batch = {**data_fn(B, g), "name": "train"}

        def one():
            for q in params:
                q.grad = None
            out = problem(batch)
            out["train_loss"].backward()
            return float(out["train_loss"])
        what = (f"UNMODIFIED reference modules (pnnl/neuromancer System/Node/MLP_bounds/RK4/PenaltyLoss/Problem from {os.path.relpath(src, ROOT) if src.startswith(ROOT) else src}, "
                f"loaded through oracle/ref_shim.py), Problem.forward + loss.backward on torch CPU fp32")
    else:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import helpers as H
        case = H.make_case(name, "cpu")
        B = CPU_PORT_BATCH.get(name, 1024)
        batch = case.make_batch(B, "cpu", 0)
        params = [p.detach().clone() for p in case.plan_params]

        def one():
            return float(H.oracle_run(name, params, batch)["loss"])
        what = "oracle port (same torch CPU ops as the reference, pinned bit-exactly to it; oracle/_ref absent)"
    t_begin = time.perf_counter()
    for _ in range(max(min(warmup, 1), 1)):
        one()
    times = []
    for _ in range(max(steps, 1)):
        t0 = time.perf_counter()
        one()
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_begin > budget_s and len(times) >= 1:
            break
    per = sorted(times)[len(times) // 2]
    return {"value": B * nsteps / per, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
            "sample": f"{name} at batch {B} (full horizon {nsteps}), median of {len(times)} timed fwd+bwd passes of the {what}",
            "ms_per_step": per * 1e3, "batch": B, "steps_timed": len(times)}


# ------------------------------------------------------------------------------------------- helpers of the B200 arm
def make_inputs(case, B, dev, seed):
    """Synthetic batch of SURVEY 8(d2) resident on `dev` (tensors a config creates with torch.zeros are made on the device)."""
    import torch
    batch = case.make_batch(B, dev, seed)
    return {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}


def quick_raw(name, dev, steps=5, warmup=3, batch=None):
    """Kernel-path throughput of another config (extra.per_config): the config's own batch, device-resident."""
    import ctypes
    import torch
    from neuromancer_b200 import _lib, configs, rollout as nr
    case = configs.make_case(name, dev)
    B = batch or case.default_batch
    data = make_inputs(case, B, dev, 0)
    plan = case.system.plan_for(case.system.init(dict(data)), loss=case.problem.loss)
    raw = nr.RawTrainStep(plan, data)
    for _ in range(warmup):
        raw.forward(); raw.backward()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        raw.forward(); raw.backward()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / steps
    info = _lib.load().nm_kernel_info(ctypes.byref(plan.c)).decode()
    out = {"batch": B, "nsteps": case.nsteps, "ms_per_step": ms, "value": B * case.nsteps / (ms * 1e-3), "unit": UNIT,
           "loss": float(raw.scalars[-1]), "kernel_info": info.split(" compiled_terms")[0],
           "note": "device-resident forward + adjoint, back to back (no L2 flush), CUDA events"}
    del raw, data, plan, case
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c5")
    ap.add_argument("--batch", type=int, default=0, help="trajectories per GPU (default: the config's)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-per-config", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    import torch
    from neuromancer_b200 import configs
    case0 = configs.build_case(args.workload)
    nsteps = case0.nsteps
    # c5 is quoted at a fixed GLOBAL batch (4M trajectories at 1/2/4/8 GPUs: strong scaling); every other config at a
    # fixed batch per GPU (weak scaling)
    strong = args.workload == "c5" and not args.batch
    per_gpu = args.batch or (case0.default_batch // max(world, 1) if strong else case0.default_batch)
    cfg = {"workload": args.workload, "nsteps": nsteps, "batch_per_gpu": per_gpu,
           "global_batch": per_gpu * max(world, 1), "parallelism": f"dp{world}",
           "description": DESCRIPTION.get(args.workload, args.workload)}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        res = time_cpu_reference(args.workload, max(min(args.steps, 5), 1), args.warmup, nsteps)
        line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": res["steps_timed"], "warmup": 1, "ms_per_step": res["ms_per_step"],
                "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": {**cfg, "cpu_sample_batch": res["batch"]},
                "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return 0

    # ------------------------------------------------------------------ B200 arm
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the fused rollout has no CPU fallback "
                         "(use --impl reference for the CPU baseline)")
    import torch.distributed as dist
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    # NCCL prints its version banner on stdout; the contract is ONE JSON line, so park stdout on stderr
    # until the result is printed
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from neuromancer_b200 import rollout as nr
    from neuromancer_b200.distributed import GradientSync

    B = cfg["batch_per_gpu"]
    case = configs.make_case(args.workload, dev)              # same seed on every rank: replicated weights
    batch = make_inputs(case, B, dev, rank)                   # per-rank shard of the sampled batch, resident in HBM
    sync = GradientSync() if world > 1 else None
    plan = case.system.plan_for(case.system.init(dict(batch)), loss=case.problem.loss)
    # data-parallel exchange: in-kernel sum over NVLink peer memory (nm_grad_allreduce: NVLS multimem.ld_reduce), NCCL only
    # when the symmetric allocation is not available on this box or NM_SYNC=nccl asks for it
    exchange = "none (one rank)"
    raw = None
    if world > 1 and os.environ.get("NM_SYNC", "peer") != "nccl":
        try:
            raw = nr.RawTrainStep(plan, batch, world_size=world, peer_sync=True)
            exchange = f"in-kernel all-reduce over symmetric memory ({raw.sync.mode}), {4 * raw.sync.n} B per step"
        except Exception as e:                                 # noqa: BLE001
            print(f"[bench] symmetric-memory exchange unavailable ({type(e).__name__}: {e}); using NCCL", file=sys.stderr)
            raw = None
    if raw is None:
        raw = nr.RawTrainStep(plan, batch, world_size=world)
        if world > 1:
            exchange = f"NCCL all-reduce, {4 * raw.gbuf.numel()} B per step"
    cfg["exchange"] = exchange
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def step_raw():
        raw.forward()
        raw.backward()
        raw.all_reduce()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        step_raw()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    launches0 = raw.launches
    barrier()
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.zero_()                                          # evict the step's working set from L2
        ev[i][0].record()
        raw.forward()
        ev[i][1].record()
        raw.backward()
        ev[i][2].record()
        raw.all_reduce()
        ev[i][3].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    if getattr(raw, "sync", None) is not None and raw.sync.status() != 0:
        raise RuntimeError("in-kernel exchange: a peer did not arrive within the kernel's time-out (status word set)")
    step_ms = [e[0].elapsed_time(e[3]) for e in ev]
    fwd_ms = sum(e[0].elapsed_time(e[1]) for e in ev) / args.steps
    bwd_ms = sum(e[1].elapsed_time(e[2]) for e in ev) / args.steps
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(total_ms) / args.steps
    launches = raw.launches - launches0
    value = world * B * nsteps / (ms_per_step * 1e-3)
    loss_val = float(raw.scalars[-1]) / world                 # the tail was SUM-all-reduced with the gradient
    import ctypes
    from neuromancer_b200 import _lib
    kernel_info = _lib.load().nm_kernel_info(ctypes.byref(plan.c)).decode()
    wide = "wide" in kernel_info
    del raw, batch, plan
    torch.cuda.empty_cache()

    # ------------------------------------------------------------------ end-to-end through the public API
    e2e = None
    if not args.no_e2e:
        Be = min(B, E2E_MAX_BATCH)
        params = [q for q in case.plan_params if q.requires_grad]
        batch_cpu = case.make_batch(Be, "cpu", rank)
        pinned = {k: (v.pin_memory() if isinstance(v, torch.Tensor) else v) for k, v in batch_cpu.items()}
        del batch_cpu
        tensors = [k for k, v in pinned.items() if isinstance(v, torch.Tensor)]
        h2d = sum(pinned[k].numel() * pinned[k].element_size() for k in tensors)

        from neuromancer_b200.dataset import DevicePrefetcher
        from neuromancer_b200.trainer import LaggedScalarReader
        reader = LaggedScalarReader(dev)                      # every step's loss is read on the host, one step late
        losses_seen = []

        def step_api(b):
            out = case.problem(b)
            for q in params:
                q.grad = None
            out["train_loss"].backward()
            if sync is not None:
                sync.sync_gradients(params)
            losses_seen.extend(reader.push(out["train_loss"]))   # device -> host read of the step's loss (4 bytes)

        def host_batches(n):                                  # every step copies its inputs from pinned host memory
            for _ in range(n):
                yield dict(pinned)
        # a step of the headline workload takes ~0.1-1 s: two warm-up steps settle the caching allocator; small
        # workloads (ms per step) keep the long warm-up of round 1 (the allocator needs ~100 steps to stop growing)
        slow = ms_per_step * (Be / B) > 50.0
        n_warm = 2 if slow else max(args.warmup, 100)
        n_timed = max(2, min(args.steps, 5)) if slow else args.steps
        n_rep = 1 if slow else 3
        for b in DevicePrefetcher(host_batches(n_warm), dev):
            step_api(b)
        losses_seen.extend(reader.drain())
        losses_seen.clear()
        import gc
        reps = []
        for _ in range(n_rep):
            gc.collect()
            gc.disable()
            barrier()
            t0 = time.perf_counter()
            for b in DevicePrefetcher(host_batches(n_timed), dev):
                step_api(b)
            losses_seen.extend(reader.drain())                # the last steps' losses: still inside the timed region
            barrier()
            dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            gc.enable()
            assert len(losses_seen) == n_timed, "every timed step's loss must have reached the host"
            losses_seen.clear()
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            reps.append(float(dt))
        dt_med = sorted(reps)[len(reps) // 2]
        e2e = {"value": world * Be * nsteps * n_timed / dt_med, "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": 4, "ms_per_step": dt_med / n_timed * 1e3, "steps": n_timed, "batch_per_gpu": Be,
               "ms_per_step_repetitions": [r / n_timed * 1e3 for r in reps],
               "api": "DevicePrefetcher (Trainer.train loop) -> Problem.forward + loss.backward; every batch copied from pinned host memory "
                      "inside the timed region (the copy of step k+1 overlaps step k), every step's loss read back on the host "
                      "(asynchronously, one step late)" + ("" if Be == B else f"; per-GPU batch capped at {Be} (host batches are double-buffered on the device)")}
        del pinned
        torch.cuda.empty_cache()

    # clocks / throttle reasons sampled over both timed regions (device-resident steps and end-to-end steps)
    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ------------------------------------------------------------------ roofline of the dominant kernel group (adjoint)
    peaks = read_peaks()
    flop_step = case.flop_per_traj_step                        # fwd+bwd algorithmic FLOP per trajectory-step
    bwd_flops = (2.0 / 3.0) * flop_step * B * nsteps          # the adjoint launches' share (2 of 3 passes)
    fwd_flops = (1.0 / 3.0) * flop_step * B * nsteps
    tf32x3_peak = peaks["bf16_tflops"] / 2.0 / 3.0             # fp32-accurate peak of a 3xTF32 split (BASELINE.md ceilings)
    fp16x3_peak = peaks["bf16_tflops"] / 3.0                   # fp32-accurate peak of the 3xFP16 split the wide kernels run
    peak = fp16x3_peak if wide else tf32x3_peak
    achieved = bwd_flops / (bwd_ms * 1e-3) / 1e12
    traffic_path = os.path.join(ROOT, "profiles", f"traffic_{args.workload}.json")
    traffic, tj = None, None
    if os.path.exists(traffic_path):                           # ncu capture of the adjoint launches, scaled to this batch
        tj = json.load(open(traffic_path))
        traffic = tj["bwd_dram_bytes_per_launch"] * (B / tj["batch"])
    if wide:
        bwd_kernel = "nm_wide_bwd_kernel + nm_wide_wgrad_kernel + term-gradient kernels (the whole adjoint window)"
    else:
        bwd_kernel = "nm_tc_bwd_kernel" if "bwd[tcgen05" in kernel_info else "nm_bwd_kernel"
    roofline = {"bound": "tensor", "kernel": bwd_kernel, "achieved": achieved, "peak": peak,
                "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_source": (f"dram read+write of the kernels named above from the committed ncu --set full capture at B = {tj['batch']} "
                                   f"(profiles/traffic_{args.workload}.json), scaled to this batch; not measured in this run") if traffic is not None else None,
                "peak_source": (f"{peaks['source']} bf16 {peaks['bf16_tflops']} TFLOP/s / 3 (3-product fp16 split, fp32-accurate)" if wide else
                                f"{peaks['source']} bf16 {peaks['bf16_tflops']} TFLOP/s / 2 (tf32) / 3 (3xTF32 fp32-accurate split)"),
                "frac_of_3xtf32_peak": achieved / tf32x3_peak, "tf32x3_peak": tf32x3_peak,
                "step_achieved_tflops": flop_step * B * nsteps / ((fwd_ms + bwd_ms) * 1e-3) / 1e12,
                "step_frac": flop_step * B * nsteps / ((fwd_ms + bwd_ms) * 1e-3) / 1e12 / peak,
                "ffma_peak_tflops": 72.2, "ffma_frac": achieved / 72.2,
                "ffma_peak_source": "measured, profiles/r01_pipe_rates_microbench.txt",
                "kernel_ms": bwd_ms, "fwd_kernel_ms": fwd_ms,
                "fwd_achieved_tflops": fwd_flops / (fwd_ms * 1e-3) / 1e12,
                "hbm_gbs_achieved": case.bytes_per_traj_step * B * nsteps / ((fwd_ms + bwd_ms) * 1e-3) / 1e9,
                "hbm_peak_gbs": peaks["hbm_gbs"]}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        res = time_cpu_reference(args.workload, 2, 1, nsteps, budget_s=30.0)
        cpu_baseline = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}

    extra = None
    if world == 1 and not args.no_per_config:
        extra = {"per_config": {}}
        for name in ("c1", "c2", "c3", "c4"):
            if name == args.workload:
                continue
            try:
                extra["per_config"][name] = quick_raw(name, dev)
            except Exception as exc:                          # a failing side config must not lose the headline line
                extra["per_config"][name] = {"error": repr(exc)[:300]}

    dtype = "f32 (policy MMAs: 3-product fp16 split, 22-bit; weight-gradient operands 1xFP16-rn)" if wide else \
            ("f32 (fwd 3xTF32, wgrad 1xTF32-rn)" if "1xTF32" in kernel_info else "f32")
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": dtype, "data": "synthetic",
            "config": {**cfg, "l2": "flushed between timed steps (256 MiB memset)", "loss": loss_val,
                       "kernel_info": kernel_info},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
            "cpu_baseline": cpu_baseline, "extra": extra, "wall_s_timed_region": t_wall}
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
