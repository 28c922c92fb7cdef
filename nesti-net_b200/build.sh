#!/bin/bash
# Builds libmups_b200.so for sm_100a (same command __graft_entry__.build() runs). Extra args go to nvcc.
set -e
cd "$(dirname "$0")"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared "$@" \
     -o libmups_b200.so csrc/*.cu
