#!/usr/bin/env bash
# Kernel-tuning experiments: builds an alternative libxfluids_b200 with extra -D flags into _variants/<name>.so
# (select it at run time with XF_LIB=...).  usage: build_variant.sh <name> "<-D flags>"
set -euo pipefail
HERE=$(cd "$(dirname "$0")" && pwd)
SRC=$HERE/csrc
name=$1; flags=${2:-}
OBJ=$HERE/_obj/var_$name
mkdir -p "$OBJ" "$HERE/_variants"
ARCH="-gencode arch=compute_100a,code=sm_100a"
COMMON="-O3 -std=c++17 -lineinfo -Xcompiler -fPIC $ARCH $flags"
nvcc $COMMON -DXF_NS=xf_strict -fmad=false -Xptxas -v -c "$SRC/xf_kernels.cu" -o "$OBJ/s.o" 2> "$OBJ/ptxas_strict.log" &
p1=$!
[ -f "$HERE/_obj/xf_kernels_fast.o" ] || { echo "run build.sh first"; exit 1; }
nvcc $COMMON -c "$SRC/xf_capi.cu" -o "$OBJ/c.o" &
p3=$!
wait $p1; wait $p3
nvcc -shared $ARCH -o "$HERE/_variants/$name.so" "$OBJ/s.o" "$HERE/_obj/xf_kernels_fast.o" "$OBJ/c.o"
python3 "$HERE/../tools/ptxas_report.py" < "$OBJ/ptxas_strict.log" | grep -E "XfCfg<5, true>, [012], 5>|k_prim<XfCfg<5|k_prim_hard<XfCfg<5" || true
echo "built _variants/$name.so"
