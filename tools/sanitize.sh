#!/bin/bash
# compute-sanitizer passes over small shapes of the hand-written kernels: memcheck, racecheck, synccheck.
# Run on a GPU box:  bash tools/sanitize.sh > gpurun_out/sanitizer.txt 2>&1
export NM_DEV_NSTEPS=6
for tool in memcheck racecheck synccheck; do
  for case in c5 c2 c3; do
    echo "=== compute-sanitizer --tool $tool   $case (B=150, nsteps=6, exact + forced tensor-core adjoint)"
    NM_DEV_CASE=$case timeout 900 compute-sanitizer --tool $tool --print-limit 5 python tools/dev_wide.py 150 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|error|hazard|B=" | cut -c1-220 | head -12
  done
done
