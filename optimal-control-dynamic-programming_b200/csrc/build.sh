#!/bin/bash
# Builds libbellman.so (sm_100a only) next to the Python package.  nvcc cross-compiles without a GPU.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../libbellman.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
ARCH="-gencode arch=compute_100a,code=sm_100a"
COMMON="-O3 -std=c++17 -lineinfo $ARCH -Xcompiler -fPIC,-ffp-contract=off,-Wall -I$HERE/../../include"
mkdir -p "$HERE/build"
# -fmad=false: one rounding per written operation (include/bellman.h); fma() calls stay fused
"$NVCC" $COMMON -fmad=false ${PTXAS_V:+-Xptxas -v} -c "$HERE/bellman_kernels.cu" -o "$HERE/build/bellman_kernels.o"
"$NVCC" $COMMON -fmad=false ${PTXAS_V:+-Xptxas -v} -c "$HERE/bellman_window.cu" -o "$HERE/build/bellman_window.o"
"$NVCC" $COMMON -fmad=false ${PTXAS_V:+-Xptxas -v} -c "$HERE/bellman_tile.cu" -o "$HERE/build/bellman_tile.o"
"$NVCC" $COMMON -c "$HERE/bellman_api.cu" -o "$HERE/build/bellman_api.o"
"$NVCC" $COMMON -x cu -c "$HERE/bellman_plan.cpp" -o "$HERE/build/bellman_plan.o"
"$NVCC" $ARCH -shared -cudart static -o "$OUT" "$HERE"/build/bellman_kernels.o "$HERE"/build/bellman_window.o "$HERE"/build/bellman_tile.o \
    "$HERE"/build/bellman_api.o "$HERE"/build/bellman_plan.o -ldl -lpthread -lrt
echo "built $OUT"
