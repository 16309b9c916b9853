#!/bin/bash
# Builds libctd_b200.so for sm_100a (cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
  -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -shared \
  -o ../libctd_b200.so engine.cu conv_tc.cu simt.cu postproc.cu segrep.cu refine.cu resize.cu "$@"
