#!/bin/bash
# A/B of library builds on one box: tools/gpu_ab.sh <rounds> <B> lib1.so lib2.so ...   (prints ms + per-phase cycles per build)
rounds=$1; B=$2; shift 2
for r in $(seq 1 $rounds); do
  for l in "$@"; do
    echo "=== round $r  $l"
    PHNN_MPC_LIB=$PWD/$l python tools/gpu_tc_phases.py $B 2>&1 | tail -24
  done
done
