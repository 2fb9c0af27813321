#!/bin/bash
for v in "" stpol1; do
  if [ -z "$v" ]; then unset NM_LIB_PATH; else export NM_LIB_PATH=neuromancer_b200/_C/variants/lib_c5_$v.so; fi
  python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu-baseline --no-per-config 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('4M variant [$v] value %.4g fwd %.3f ms  adjoint %.3f ms  sm %s' % (d['value'], d['roofline']['fwd_kernel_ms'], d['roofline']['kernel_ms'], d['clocks']['sm_mhz']))"
done
