#!/bin/bash
# Builds libbellman.so (sm_100a only) next to the Python package.  nvcc cross-compiles without a GPU.
# Translation units are compiled in parallel and only when their object is older than the source or
# any header (FORCE=1 rebuilds everything).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../libbellman.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
ARCH="-gencode arch=compute_100a,code=sm_100a"
COMMON="-O3 -std=c++17 -lineinfo $ARCH -Xcompiler -fPIC,-ffp-contract=off,-Wall -I$HERE/../../include"
mkdir -p "$HERE/build"

stale() {   # stale <object> <source>: true when the object must be rebuilt
    local obj="$1" src="$2"
    [[ -n "${FORCE:-}" || ! -f "$obj" || "$src" -nt "$obj" || "$HERE/build.sh" -nt "$obj" ]] && return 0
    local hdr
    for hdr in "$HERE"/*.h "$HERE"/*.cuh "$HERE/../../include/bellman.h"; do
        [[ "$hdr" -nt "$obj" ]] && return 0
    done
    return 1
}

pids=()
compile() {   # compile <name> <extra flags...>
    local name="$1"; shift
    local src="$HERE/$name" obj="$HERE/build/${name%.*}.o"
    if stale "$obj" "$src"; then
        "$NVCC" $COMMON "$@" ${PTXAS_V:+-Xptxas -v} -c "$src" -o "$obj" &
        pids+=($!)
    fi
}
# -fmad=false: one rounding per written operation (include/bellman.h); fma() calls stay fused
compile bellman_window.cu -fmad=false
compile bellman_kernels.cu -fmad=false
compile bellman_tile.cu -fmad=false
compile bellman_stream.cu -fmad=false
compile bellman_dense6.cu -fmad=false
compile bellman_api.cu
compile bellman_plan.cpp -x cu
for p in "${pids[@]:-}"; do [[ -n "$p" ]] && wait "$p"; done
"$NVCC" $ARCH -shared -cudart static -o "$OUT" "$HERE"/build/bellman_kernels.o "$HERE"/build/bellman_window.o \
    "$HERE"/build/bellman_tile.o "$HERE"/build/bellman_stream.o "$HERE"/build/bellman_dense6.o \
    "$HERE"/build/bellman_api.o "$HERE"/build/bellman_plan.o -ldl -lpthread -lrt
echo "built $OUT"
