"""Where does the public-API step spend host time?  (run on the GPU box)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch, helpers as H
dev = torch.device("cuda:0")
case = H.make_case("c2", dev)
B = 65536
batch_cpu = case.make_batch(B, "cpu", 0)
pinned = {k: (v.pin_memory() if isinstance(v, torch.Tensor) else v) for k, v in batch_cpu.items()}
params = [q for q in case.plan_params if q.requires_grad]
def sync(): torch.cuda.synchronize()
def step(timers):
    t0 = time.perf_counter()
    b = {k: (v.to(dev, non_blocking=True) if isinstance(v, torch.Tensor) else v) for k, v in pinned.items()}
    t1 = time.perf_counter()
    out = case.problem(b)
    t2 = time.perf_counter()
    for q in params: q.grad = None
    out["train_loss"].backward()
    t3 = time.perf_counter()
    v = float(out["train_loss"].detach())
    t4 = time.perf_counter()
    for k, d in zip(("h2d_issue", "forward_call", "backward_call", "loss_read"), (t1 - t0, t2 - t1, t3 - t2, t4 - t3)):
        timers[k] = timers.get(k, 0) + d
for _ in range(10): step({})
sync()
for rep in range(3):
    timers = {}; t0 = time.perf_counter()
    for _ in range(50): step(timers)
    sync(); tot = time.perf_counter() - t0
    print(f"rep{rep}: {tot/50*1e3:.3f} ms/step  " + "  ".join(f"{k} {v/50*1e3:.3f}" for k, v in timers.items()))
# pure H2D
sync(); t0 = time.perf_counter()
for _ in range(50):
    b = {k: (v.to(dev, non_blocking=True) if isinstance(v, torch.Tensor) else v) for k, v in pinned.items()}
sync(); print(f"h2d only: {(time.perf_counter()-t0)/50*1e3:.3f} ms  ({sum(v.numel()*4 for v in pinned.values() if isinstance(v, torch.Tensor))/1e6:.1f} MB)")
