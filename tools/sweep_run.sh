#!/bin/bash
# Run bench.py once per tuning-variant library (built by tools/sweep_build.py) and print one line each.
wl=$1
for lib in neuromancer_b200/_C/variants/lib_${wl}_*.so; do
  NM_LIB_PATH=$PWD/$lib python bench.py --workload $wl --steps 3 --warmup 2 --batch ${NM_SWEEP_BATCH:-0} --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); r=d['roofline']
print('$lib'.split('lib_')[-1], '| traj-steps/s %.3e | step %.3f ms fwd %.3f bwd %.3f | %s' % (d['value'], d['ms_per_step'], r['fwd_kernel_ms'], r['kernel_ms'], d['config']['kernel_info']))"
done
