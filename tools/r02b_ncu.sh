#!/bin/bash
# --set full capture of the wide kernels at one tile pair per SM (B = 37,888)
set -x
python bench.py --workload c5 --batch 37888 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-per-config > gpurun_out/r02b_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'nm_wide_(fwd|bwd)' -s 6 -c 2 -f -o gpurun_out/r02b_prof_c5_wide \
    python bench.py --workload c5 --batch 37888 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-per-config > gpurun_out/r02b_ncu_c5.log 2>&1
tail -3 gpurun_out/r02b_ncu_c5.log
