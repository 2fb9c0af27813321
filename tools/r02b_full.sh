#!/bin/bash
# full GPU suite + the C5 bench line (kernel path only)
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02b_gpu_tests.txt 2>&1; tail -4 gpurun_out/r02b_gpu_tests.txt
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02b_bench_c5.json 2> gpurun_out/r02b_bench_c5.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02b_bench_c5.json").read().strip().splitlines()[-1])
print("C5 value %.4g ms/step %.1f fwd %.1f adj %.1f clocks %s" % (d["value"], d["ms_per_step"], d["roofline"]["fwd_kernel_ms"], d["roofline"]["kernel_ms"], d["clocks"]))
for k,v in (d.get("extra") or {}).get("per_config",{}).items(): print(k, "%.4g" % v["value"], v["ms_per_step"])
PY
