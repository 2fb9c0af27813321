#!/bin/bash
# Round-2 evidence run (one GPU).  Outputs under gpurun_out/; summarised into profiles/ by hand (profiles/README.md).
set -x
# 1. the default bench command (C5, global batch 4,194,304), plain, then its launch list
/usr/bin/time -v python bench.py > gpurun_out/r02_bench_c5_default.json 2> gpurun_out/r02_bench_c5_default.err
python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-per-config > gpurun_out/r02_plain_a.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_c5_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-per-config > gpurun_out/r02_ncu_launches.log 2>&1
# 2. --set full capture of the wide kernels (one tile pair per SM: B = 37,888; same kernels, shorter replay)
python bench.py --workload c5 --batch 37888 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-per-config > gpurun_out/r02_plain_b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'nm_wide_(fwd|bwd|wgrad|termgrad|score)' -s 12 -c 6 -f -o gpurun_out/r02_prof_c5_wide \
    python bench.py --workload c5 --batch 37888 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-per-config > gpurun_out/r02_ncu_c5.log 2>&1
# 3. other workloads, one bench line each
for w in c1 c2 c3 c4; do python bench.py --workload $w --steps 20 --warmup 5 --no-per-config > gpurun_out/r02_bench_$w.json 2> gpurun_out/r02_bench_$w.err; done
# 4. the reference arm as the driver runs it
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err
