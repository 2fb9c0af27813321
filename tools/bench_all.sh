#!/bin/bash
# bench every workload once (kernel path only) and print a one-line summary each
for spec in "c1 3333" "c2 65536" "c3 262144" "c4 131072" "c4p 131072" "c5 16384"; do
  set -- $spec
  python bench.py --workload $1 --batch $2 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); r=d['roofline']
print('$1 B=$2 | traj-steps/s %.3e | step %.3f ms fwd %.3f bwd %.3f | bwd %.2f TFLOP/s (%.1f%% ffma) | %s' % (d['value'], d['ms_per_step'], r['fwd_kernel_ms'], r['kernel_ms'], r['achieved'], 100*r['ffma_frac'], d['config']['kernel_info']))"
done
