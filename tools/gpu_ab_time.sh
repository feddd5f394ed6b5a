#!/bin/bash
# A/B timing of library builds on one box: tools/gpu_ab_time.sh <rounds> "<workloads>" lib1.so lib2.so ...
rounds=$1; wls=$2; shift 2
for r in $(seq 1 $rounds); do
  for l in "$@"; do
    PHNN_MPC_LIB=$PWD/$l python tools/gpu_time_lib.py $wls 2>&1 | tail -n 4
  done
done
