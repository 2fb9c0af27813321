#!/bin/bash
# compute-sanitizer passes over small shapes of the hand-written kernels: memcheck, racecheck, synccheck.
# NOTE: compute-sanitizer is CLOSED on the graft GPU pool (it exits with code 86 there); kept for boxes where it runs.
# Run on a GPU box:  bash tools/sanitize.sh > gpurun_out/r02_sanitizer.txt 2>&1
# Every run prints the parity line of tools/dev_wide.py (exact adjoint, then the forced tensor-core adjoint) and the
# sanitizer's own summary line; a run that printed no summary did not finish and counts as failed.
export NM_DEV_NSTEPS=6
for tool in memcheck racecheck synccheck; do
  for case in ${NM_SANITIZE_CASES:-c5 c4 c2 c3}; do
    echo "=== compute-sanitizer --tool $tool   $case (B=150, nsteps=6, exact + forced tensor-core adjoint)"
    NM_DEV_CASE=$case timeout 1200 compute-sanitizer --tool $tool --print-limit 5 python tools/dev_wide.py 150 > /tmp/san.log 2>&1
    echo "exit code $?"
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Invalid|hazard|B=150" /tmp/san.log | cut -c1-200 | head -12
  done
done
