#!/bin/bash
# Build libgci_cuda.so in-tree for sm_100a (the .so is git-ignored but travels to the GPU box).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../libgci_cuda.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O3,-Wall,-Wno-unknown-pragmas -shared
       -cudart static ${GCI_NVCC_EXTRA:-})
"$NVCC" "${FLAGS[@]}" "$@" -o "$OUT" "$HERE/api.cu" "$HERE/filter.cu" "$HERE/depth.cu" "$HERE/scan.cu" "$HERE/gzip.cu" "$HERE/comm.cu" "$HERE/shard.cu" -ldl
echo "built $OUT"
CXX="${CXX:-g++}"
"$CXX" -O3 -std=c++17 -fPIC -shared -Wall -pthread -o "$HERE/../libgci_io.so" "$HERE/io_native.cpp" -lz
echo "built $HERE/../libgci_io.so"
