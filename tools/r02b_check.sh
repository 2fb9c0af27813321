#!/bin/bash
# Quick GPU check of a wide-kernel change: parity tests of the wide path, the C5 bench line (kernel path only), phase profile.
set -x
python -m pytest tests/test_wide_gpu.py -m gpu -x -q > gpurun_out/r02b_wide_tests.txt 2>&1; tail -3 gpurun_out/r02b_wide_tests.txt
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-per-config > gpurun_out/r02b_bench_c5.json 2> gpurun_out/r02b_bench_c5.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02b_bench_c5.json").read().strip().splitlines()[-1])
print("C5 value %.4g ms/step %.1f fwd %.1f adj %.1f clocks %s" % (d["value"], d["ms_per_step"], d["roofline"]["fwd_kernel_ms"], d["roofline"]["kernel_ms"], d["clocks"]))
PY
NM_LIB_PATH=neuromancer_b200/_C/variants/lib_c5_wprof.so python bench.py --workload c5 --batch 37888 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-per-config 2>&1 | grep wprof | tail -9 > gpurun_out/r02b_wprof_c5.txt; cat gpurun_out/r02b_wprof_c5.txt
