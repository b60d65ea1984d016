#!/usr/bin/env bash
# Builds xfluids_b200/libxfluids_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
# The kernel source is compiled twice: strict (-fmad=false, parity mode) and fast (FMA contraction allowed).
# Objects are rebuilt only when one of their sources is newer (XF_FORCE_BUILD=1: everything).
set -euo pipefail
HERE=$(cd "$(dirname "$0")" && pwd)
SRC=$HERE/csrc
OBJ=$HERE/_obj
INC=$HERE/../include/xfluids_b200.h
mkdir -p "$OBJ"
ARCH="-gencode arch=compute_100a,code=sm_100a"
COMMON="-O3 -std=c++17 -lineinfo -Xcompiler -fPIC $ARCH ${XF_EXTRA:-}"
FORCE=${XF_FORCE_BUILD:-0}
stale() { # stale <target> <deps...>
  local t=$1; shift
  [ "$FORCE" = 1 ] && return 0
  [ -f "$t" ] || return 0
  for d in "$@"; do [ "$d" -nt "$t" ] && return 0; done
  return 1
}
KDEPS="$SRC/xf_kernels.cu $SRC/xf_math.cuh $SRC/xf_march.cuh $SRC/xf_visc.cuh $SRC/xf_tma.cuh $SRC/xf_log.cuh $SRC/xf_log_data.h $SRC/xf_types.h $SRC/xf_launch.h"
pids=()
if stale "$OBJ/xf_kernels_strict.o" $KDEPS; then nvcc $COMMON -DXF_NS=xf_strict -fmad=false -c "$SRC/xf_kernels.cu" -o "$OBJ/xf_kernels_strict.o" & pids+=($!); fi
if stale "$OBJ/xf_kernels_fast.o" $KDEPS; then nvcc $COMMON -DXF_NS=xf_fast -fmad=true -c "$SRC/xf_kernels.cu" -o "$OBJ/xf_kernels_fast.o" & pids+=($!); fi
if stale "$OBJ/xf_capi.o" "$SRC/xf_capi.cu" "$SRC/xf_launch.h" "$SRC/xf_types.h" "$SRC/xf_log.cuh" "$SRC/xf_log_data.h" "$INC"; then nvcc $COMMON -c "$SRC/xf_capi.cu" -o "$OBJ/xf_capi.o" & pids+=($!); fi
if stale "$OBJ/xf_slab.o" "$SRC/xf_slab.cu" "$INC"; then nvcc $COMMON -c "$SRC/xf_slab.cu" -o "$OBJ/xf_slab.o" & pids+=($!); fi
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
if stale "$HERE/libxfluids_b200.so" "$OBJ/xf_kernels_strict.o" "$OBJ/xf_kernels_fast.o" "$OBJ/xf_capi.o" "$OBJ/xf_slab.o"; then
  nvcc -shared $ARCH -o "$HERE/libxfluids_b200.so" "$OBJ/xf_kernels_strict.o" "$OBJ/xf_kernels_fast.o" "$OBJ/xf_capi.o" "$OBJ/xf_slab.o" -ldl
  echo "built $HERE/libxfluids_b200.so"
else
  echo "libxfluids_b200.so up to date"
fi
# host layer (C++17, no CUDA): Setup / initial conditions / XFLUIDS driver.  -ffp-contract=off: derived metrics and
# initial states must round exactly like the reference's parity build.
H=$HERE/host
HF="-std=c++17 -O2 -ffp-contract=off -fopenmp -fPIC -pthread"
HSRC=$(ls "$H"/xfh_*.cpp)
if stale "$HERE/libxfluids_host.so" $HSRC "$H"/*.hpp "$INC" "$HERE/libxfluids_b200.so" || stale "$HERE/xfluids" "$H/main.cpp" "$HERE/libxfluids_host.so"; then
  g++ $HF -shared $HSRC -o "$HERE/libxfluids_host.so" -L"$HERE" -lxfluids_b200 -Wl,-rpath,'$ORIGIN'
  g++ $HF "$H/main.cpp" -o "$HERE/xfluids" -L"$HERE" -lxfluids_host -lxfluids_b200 -Wl,-rpath,'$ORIGIN'
  echo "built $HERE/libxfluids_host.so and $HERE/xfluids"
fi
