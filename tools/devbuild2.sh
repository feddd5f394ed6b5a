#!/bin/bash
# fast experiment build: only the pendulum shape (pHNN learned G, n=2, h=64); usage: tools/devbuild2.sh out.so [extra nvcc flags]
out=$1; shift
cd "$(dirname "$0")/../phnn_mpc_b200/csrc" && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --shared -Xcompiler -fPIC \
  -DPHNN_DEV_CFG2 \
  -Xptxas -v "$@" -o "$out" phnn_capi.cu 2>&1 | grep -E "error|registers|spill" | grep -v "^$" | sort | uniq -c | sort -rn | head -8
