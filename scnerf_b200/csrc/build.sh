#!/bin/bash
# Build libscnerf_b200.so for sm_100a (in-tree; the .so is git-ignored but travels with gpurun).
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
  -Xcompiler -fPIC -shared -o libscnerf_b200.so api.cu "$@"
echo "built $(pwd)/libscnerf_b200.so"
