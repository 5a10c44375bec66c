#!/bin/bash
# Build libscnerf_b200.so for sm_100a (in-tree; the .so is git-ignored but travels with gpurun).
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
# OUT=libscnerf_b200_timeline.so bash build.sh -DSCNERF_TIMELINE   builds the variant with the in-kernel timeline stamps
# (tools/timeline_pipe.py, run with SCNERF_LIB=<that file>); the product build carries none.
OUT=${OUT:-libscnerf_b200.so}
# 20013: the slab tables are built by constexpr functions that also take host lambdas (pack tables): benign, very noisy
$NVCC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -diag-suppress 20013 \
  -Xcompiler -fPIC -shared -o $OUT api.cu "$@"
echo "built $(pwd)/$OUT"
