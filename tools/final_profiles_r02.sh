#!/bin/bash
# Round-2 evidence run (one GPU).  Outputs under gpurun_out/; tools/summarize_r02.sh turns them into profiles/.
set -x
# 1. the default bench command (C5, global batch 4,194,304), plain, then its launch list
python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_c5_default.json 2> gpurun_out/r02_bench_c5_default.err
python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-per-config > gpurun_out/r02_plain_a.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_c5_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-per-config > gpurun_out/r02_ncu_launches.log 2>&1
# 2. --set full capture of the wide kernels (one tile pair per SM: B = 37,888; same kernels, shorter replay)
python bench.py --workload c5 --batch 37888 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-per-config > gpurun_out/r02_plain_b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'nm_wide_(fwd|bwd|wgrad|termgrad|score)' -s 12 -c 6 -f -o gpurun_out/r02_prof_c5_wide \
    python bench.py --workload c5 --batch 37888 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-per-config > gpurun_out/r02_ncu_c5.log 2>&1
# 3. c4 with the first-layer hoist
python bench.py --workload c4 --steps 10 --warmup 5 --no-per-config > gpurun_out/r02_bench_c4.json 2> gpurun_out/r02_bench_c4.err
ncu --set full --clock-control none --import-source on -k regex:'nm_(tc_fwd|tc_bwd|exo_gemm)' -s 16 -c 4 -f -o gpurun_out/r02_prof_c4_hoist \
    python bench.py --workload c4 --batch 32768 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-per-config > gpurun_out/r02_ncu_c4.log 2>&1
# 4. phase profile of the adjoint (clock64 instrumentation, variant library)
NM_LIB_PATH=neuromancer_b200/_C/variants/lib_c5_wprof.so python bench.py --workload c5 --batch 37888 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-per-config 2>&1 | grep wprof | tail -3 > gpurun_out/r02_wprof_c5.txt
# 5. other workloads, one bench line each
for w in c1 c2 c3; do python bench.py --workload $w --steps 20 --warmup 5 --no-per-config > gpurun_out/r02_bench_$w.json 2> gpurun_out/r02_bench_$w.err; done
