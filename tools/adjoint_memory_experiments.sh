#!/bin/bash
# Timing experiments behind profiles/r02_adjoint_memory_experiments.txt and profiles/r02_l2_carveout_sweep.txt (one B200).
# The variant libraries leave out / redirect memory traffic of the wide adjoint -- THEIR RESULTS ARE WRONG ON PURPOSE, they
# only answer "what would the sweep cost without ...":
#   wnostage=1    no staging stores            wnostash=1   no derivative stash traffic
#   wstwrap=128   staging wrapped into a 34 MB (L2-resident) window          wstpol=1|2  staging stores default / .cg instead of .cs
# Build them here (build container), run on the GPU box:   gpurun -- 'bash tools/adjoint_memory_experiments.sh run'
if [ "$1" != "run" ]; then
  python - <<'PY'
from neuromancer_b200 import build
for tag, tune in (("nostage", "wnostage=1"), ("nostash", "wnostash=1"), ("nostash_nostage", "wnostash=1,wnostage=1"),
                  ("wrap128", "wstwrap=128"), ("stpol1", "wstpol=1"), ("stpol2", "wstpol=2")):
    print(build.build_variant("c5", tune, f"neuromancer_b200/_C/variants/lib_c5_{tag}.so"))
PY
  exit 0
fi
run() {  # $1 = batch, env NM_LIB_PATH / NM_WIDE_L2_PERSIST_MB select the variant
  python bench.py --workload c5 --batch $1 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-per-config 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('B $1 variant [${NM_LIB_PATH##*/}] carve-out [${NM_WIDE_L2_PERSIST_MB:-default}] fwd %.3f ms  adjoint window %.3f ms  sm %s MHz' % (d['roofline']['fwd_kernel_ms'], d['roofline']['kernel_ms'], d['clocks']['sm_mhz']))"
}
export NM_TC_BWD=1
for b in 128 256 18944 37888; do
  for v in "" nostage nostash nostash_nostage wrap128 stpol1 stpol2; do
    if [ -z "$v" ]; then unset NM_LIB_PATH; else export NM_LIB_PATH=neuromancer_b200/_C/variants/lib_c5_$v.so; fi
    run $b
  done
done
unset NM_LIB_PATH NM_TC_BWD
for mb in 0 16 32 48 56 64 72 88 126; do NM_WIDE_L2_PERSIST_MB=$mb run 4194304; done
