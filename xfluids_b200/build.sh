#!/usr/bin/env bash
# Builds xfluids_b200/libxfluids_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
# The kernel source is compiled twice: strict (-fmad=false, parity mode) and fast (FMA contraction allowed).
set -euo pipefail
HERE=$(cd "$(dirname "$0")" && pwd)
SRC=$HERE/csrc
OBJ=$HERE/_obj
mkdir -p "$OBJ"
ARCH="-gencode arch=compute_100a,code=sm_100a"
COMMON="-O3 -std=c++17 -lineinfo -Xcompiler -fPIC $ARCH"
up_to_date() { [ -f "$1" ] && [ -z "$(find "$SRC" "$HERE/../include" -newer "$1" -type f | head -1)" ]; }
if up_to_date "$HERE/libxfluids_b200.so" && [ "${XF_FORCE_BUILD:-0}" != 1 ]; then echo "libxfluids_b200.so up to date"; exit 0; fi
nvcc $COMMON -DXF_NS=xf_strict -fmad=false -c "$SRC/xf_kernels.cu" -o "$OBJ/xf_kernels_strict.o" &
p1=$!
nvcc $COMMON -DXF_NS=xf_fast -fmad=true -c "$SRC/xf_kernels.cu" -o "$OBJ/xf_kernels_fast.o" &
p2=$!
nvcc $COMMON -c "$SRC/xf_capi.cu" -o "$OBJ/xf_capi.o" &
p3=$!
wait $p1; wait $p2; wait $p3
nvcc -shared $ARCH -o "$HERE/libxfluids_b200.so" "$OBJ/xf_kernels_strict.o" "$OBJ/xf_kernels_fast.o" "$OBJ/xf_capi.o"
echo "built $HERE/libxfluids_b200.so"
