#!/usr/bin/env bash
# Builds xfluids_b200/libxfluids_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
# The kernel source is compiled twice: strict (-fmad=false, parity mode) and fast (FMA contraction allowed).
set -euo pipefail
HERE=$(cd "$(dirname "$0")" && pwd)
SRC=$HERE/csrc
OBJ=$HERE/_obj
mkdir -p "$OBJ"
ARCH="-gencode arch=compute_100a,code=sm_100a"
COMMON="-O3 -std=c++17 -lineinfo -Xcompiler -fPIC $ARCH ${XF_EXTRA:-}"
up_to_date() { [ -f "$1" ] && [ -z "$(find "$SRC" "$HERE/../include" -newer "$1" -type f | head -1)" ]; }
host_up_to_date() { [ -f "$1" ] && [ -z "$(find "$HERE/host" "$HERE/../include" "$HERE/libxfluids_b200.so" -newer "$1" -type f | head -1)" ]; }
build_host() {
  # host layer (C++17, no CUDA): Setup / initial conditions / XFLUIDS driver.  -ffp-contract=off: derived metrics and
  # initial states must round exactly like the reference's parity build.
  local H=$HERE/host
  local HF="-std=c++17 -O2 -ffp-contract=off -fopenmp -fPIC"
  g++ $HF -shared "$H/xfh_setup.cpp" "$H/xfh_ini.cpp" "$H/xfh_driver.cpp" "$H/xfh_capi.cpp" -o "$HERE/libxfluids_host.so" -L"$HERE" -lxfluids_b200 -Wl,-rpath,'$ORIGIN'
  g++ $HF "$H/main.cpp" -o "$HERE/xfluids" -L"$HERE" -lxfluids_host -lxfluids_b200 -Wl,-rpath,'$ORIGIN'
  echo "built $HERE/libxfluids_host.so and $HERE/xfluids"
}
if up_to_date "$HERE/libxfluids_b200.so" && [ "${XF_FORCE_BUILD:-0}" != 1 ]; then
  echo "libxfluids_b200.so up to date"
  if ! host_up_to_date "$HERE/libxfluids_host.so" || [ ! -x "$HERE/xfluids" ]; then build_host; fi
  exit 0
fi
nvcc $COMMON -DXF_NS=xf_strict -fmad=false -c "$SRC/xf_kernels.cu" -o "$OBJ/xf_kernels_strict.o" &
p1=$!
nvcc $COMMON -DXF_NS=xf_fast -fmad=true -c "$SRC/xf_kernels.cu" -o "$OBJ/xf_kernels_fast.o" &
p2=$!
nvcc $COMMON -c "$SRC/xf_capi.cu" -o "$OBJ/xf_capi.o" &
p3=$!
wait $p1; wait $p2; wait $p3
nvcc -shared $ARCH -o "$HERE/libxfluids_b200.so" "$OBJ/xf_kernels_strict.o" "$OBJ/xf_kernels_fast.o" "$OBJ/xf_capi.o"
echo "built $HERE/libxfluids_b200.so"
build_host
