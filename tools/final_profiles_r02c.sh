#!/bin/bash
# Round-2 (final session) evidence run, one GPU.  Outputs under gpurun_out/; summarised into profiles/ (profiles/README.md).
set -x
python -m pytest tests -m gpu -q > gpurun_out/r02c_gpu_tests.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02c_smoke.txt 2>&1
# the default bench command (C5, global batch 4,194,304), plain, then its launch list
python bench.py > gpurun_out/r02c_bench_c5_default.json 2> gpurun_out/r02c_bench_c5_default.err
python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-per-config > gpurun_out/r02c_plain_a.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02c_c5_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-per-config > gpurun_out/r02c_ncu_launches.log 2>&1
# other workloads, one bench line each
for w in c1 c2 c3 c4; do python bench.py --workload $w --steps 20 --warmup 5 --no-per-config > gpurun_out/r02c_bench_$w.json 2> gpurun_out/r02c_bench_$w.err; done
# the reference arm as the driver runs it
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02c_bench_reference_arm.json 2> gpurun_out/r02c_bench_reference_arm.err
tail -3 gpurun_out/r02c_gpu_tests.txt; tail -2 gpurun_out/r02c_smoke.txt
