#!/usr/bin/env bash
# Tuning builds: xfluids_b200/variants/lib_<name>.so = the library with extra -D flags, only the SBI (Emax = 9, WENO5, no limiter)
# kernels instantiated.  Selected at run time with XF_LIB=<path> (xfluids_b200/capi.py).   usage: tools/build_variant.sh <name> [-DFOO=1 ...]
set -euo pipefail
HERE=$(cd "$(dirname "$0")/.." && pwd)/xfluids_b200
NAME=$1; shift
OBJ=$HERE/_obj/var_$NAME; mkdir -p "$OBJ" "$HERE/variants"
ARCH="-gencode arch=compute_100a,code=sm_100a"
COMMON="-O3 -std=c++17 -lineinfo -Xcompiler -fPIC $ARCH -DXF_ONLY_SBI $*"
nvcc $COMMON -DXF_NS=xf_strict -fmad=false -Xptxas -v -c "$HERE/csrc/xf_kernels.cu" -o "$OBJ/strict.o" 2> "$OBJ/ptxas_strict.log" &
nvcc $COMMON -DXF_NS=xf_fast -fmad=true -c "$HERE/csrc/xf_kernels.cu" -o "$OBJ/fast.o" 2> /dev/null &
nvcc $COMMON -c "$HERE/csrc/xf_capi.cu" -o "$OBJ/capi.o" &
nvcc $COMMON -c "$HERE/csrc/xf_slab.cu" -o "$OBJ/slab.o" &
wait
nvcc -shared $ARCH -o "$HERE/variants/lib_$NAME.so" "$OBJ/strict.o" "$OBJ/fast.o" "$OBJ/capi.o" "$OBJ/slab.o" -ldl
grep -A2 "k_march" "$OBJ/ptxas_strict.log" | grep -E "spill|registers" | paste - - | sed "s/^/$NAME: /"
