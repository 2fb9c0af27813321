"""Developer script: run every case on the GPU and print the deviation from the CPU oracle."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import helpers as H

names = sys.argv[1:] or list(H.configs.CASES)
dev = "cuda:0"
for name in names:
    B = {"c5": 96, "c4": 200, "c3": 300}.get(name, 333)
    t0 = time.time()
    try:
        case = H.make_case(name, dev)
        batch_cpu = case.make_batch(B, "cpu", 1)
        batch = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in batch_cpu.items()}
        params_cpu = [p.detach().cpu().clone() for p in case.plan_params]
        got = H.product_run(case, batch)
        torch.cuda.synchronize()
        ref = H.oracle_run(name, params_cpu, batch_cpu)
        ref64 = H.oracle_run(name, params_cpu, batch_cpu, dtype=torch.float64)
        line = [f"{name:24s} B={B}"]
        for k in got["traj"]:
            line.append(f"{k}:{H.rel_err(got['traj'][k], ref['traj'][k]):.2e}/{H.rel_err(got['traj'][k], ref64['traj'][k]):.2e}(ref32-64 {H.rel_err(ref['traj'][k], ref64['traj'][k]):.2e})")
        line.append(f"loss:{H.rel_err(got['loss'], ref['loss']):.2e} [{float(got['loss']):.6g} vs {float(ref['loss']):.6g}]")
        gerr = max(H.rel_err(g, r) for g, r in zip(got["grads"], ref["grads"]))
        gerr64 = max(H.rel_err(g, r) for g, r in zip(got["grads"], ref64["grads"]))
        r3264 = max(H.rel_err(g, r) for g, r in zip(ref["grads"], ref64["grads"]))
        line.append(f"grads:{gerr:.2e}/{gerr64:.2e}(ref32-64 {r3264:.2e})")
        terr = max(H.rel_err(got["terms"][k], ref["terms"][k]) for k in got["terms"])
        line.append(f"terms:{terr:.2e}")
        print("  ".join(line), f"[{time.time()-t0:.1f}s]", flush=True)
    except Exception as e:
        import traceback; traceback.print_exc()
        print(f"{name}: FAILED {type(e).__name__}: {e}", flush=True)
