#!/bin/bash
# Round-end evidence run (one GPU): bench every workload, launch list of the default bench command, and one
# `ncu --set full` capture each of the kernels of c2 (tcgen05, policy mode), c3 (tcgen05, RHS mode) and c4 (FFMA).  Outputs under gpurun_out/.
set -x
bash tools/bench_all.sh > gpurun_out/bench_all.txt 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'nm_tc_(fwd|bwd)_kernel' -s 6 -c 2 -f -o gpurun_out/prof_c2_final \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_c2.log 2>&1
python bench.py --workload c3 --batch 65536 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'nm_tc_(fwd|bwd)_kernel' -s 6 -c 2 -f -o gpurun_out/prof_c3_final \
    python bench.py --workload c3 --batch 65536 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_c3.log 2>&1
python bench.py --workload c4 --batch 32768 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/plain4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'nm_(fwd|bwd)_kernel' -s 6 -c 2 -f -o gpurun_out/prof_c4_final \
    python bench.py --workload c4 --batch 32768 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_c4.log 2>&1
