#!/bin/bash
# strong scaling of the default bench command (C5, global batch 4,194,304) at N GPUs of one box, launched the way the driver does
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --no-per-config > gpurun_out/r02c_scale_n$N.json 2> gpurun_out/r02c_scale_n$N.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r02c_scale_n$N.json") if l.startswith("{")][-1])
e=d.get("e2e") or {}
print("N=%d %s B_global %s | traj-steps/s %.4g | step %.1f ms | fwd %.1f adj %.1f | e2e %.4g | clocks %s" % (d["n_gpus"], d["config"]["workload"], d["config"].get("global_batch"), d["value"], d["ms_per_step"], d["roofline"]["fwd_kernel_ms"], d["roofline"]["kernel_ms"], e.get("value") or 0, d["clocks"]))
PY
