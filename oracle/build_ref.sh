#!/usr/bin/env bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE.
# Compiles the UNMODIFIED XFluids reference sources where they lie under /root/reference
# (nothing is copied into this repo) against the host SYCL shim in oracle/shim, once per
# case/scheme, into oracle/_ref/<case>_w<order>_<mode>/XFLUIDS.   SURVEY.md 8c / Appendix E.
#
#   usage: oracle/build_ref.sh <case> <weno 5|6|7> <mode parity|fast> [alpha LLF|GLF|ROE] [pp 0|1] [visc 0|1]
#   out 1 (7th) = oracle/cases/<case>_out.json: OutDAT / OutVTI on and one output stamp per format (common / compressed / partial)
#   visc 1 = -DVisc=1 -DVisc_Heat=1 -DVisc_Diffu=1 with the fourth-order viscous discretisation (what init_sample.cmake sets for shock-bubble)
#   weno 6 = WENO-CU6 (SCHEME_ORDER 6); pp 1 = oracle/cases/<case>_pp.json: equations.PositivityPreserving true at CFL 0.9 (at the cases' CFL 0.4 the limiter
#   never acts in these flows; at 0.9 it limits from the first step on)
#   case:  shock-tube | vortex | riemann | sbi | jet
#   parity: -O2 -ffp-contract=off, serial  (the bit-level oracle)
#   fast:   -O3 -march=native -fopenmp     (the CPU throughput baseline)
#
# The macro set is what cmake/init_sample.cmake + cmake/init_options.cmake would define for
# the case, with the north_star overrides (inviscid, LLF, WENO5/7, PP off, reactions off).
set -euo pipefail
REF=${XF_REFERENCE:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
REPO=$(cd "$HERE/.." && pwd)
CASE=$1; WENO=${2:-5}; MODE=${3:-parity}; ALPHA=${4:-LLF}; PP=${5:-0}; VISC=${6:-0}; OUTJ=${7:-0}
[ -d "$REF/src" ] || { echo "reference tree $REF not present: cannot build oracle/_ref (prebuilt files are used on the GPU box)"; exit 3; }

case $CASE in
  shock-tube) SAMPLE=1d-insert-st;           SDIR=src/solver_Ini/sample/1D-X-Y-Z/insert-st;         MIX=1d-mc-insert-shock-tube; COP=1;;
  vortex)     SAMPLE=2d-euler-vortex;        SDIR=src/solver_Ini/sample/2D-EulerVortex;              MIX=NO-COP;                  COP=0;;
  riemann)    SAMPLE=2d-riemann-shocks;      SDIR=src/solver_Ini/sample/2D-Riemann/shocks-interaction; MIX=NO-COP;                COP=0;;
  sbi)        SAMPLE=shock-bubble;           SDIR=src/solver_Ini/sample/shock-bubble-intera;         MIX=Inert-SBI;               COP=1;;
  jet)        SAMPLE=3d-under-expanded-jet;  SDIR=src/solver_Ini/sample/under-expanded-jet;          MIX=2d-under-expanded-jet;   COP=1;;
  *) echo "unknown case $CASE"; exit 2;;
esac
case $ALPHA in ROE) AT=1;; LLF) AT=2;; GLF) AT=3;; *) echo "bad alpha"; exit 2;; esac

TAG=${CASE}_w${WENO}_${MODE}; [ "$ALPHA" != LLF ] && TAG=${TAG}_${ALPHA}; [ "$PP" = 1 ] && TAG=${TAG}_pp; [ "$VISC" = 1 ] && TAG=${TAG}_visc; [ "$OUTJ" = 1 ] && TAG=${TAG}_out
OUT=$HERE/_ref/$TAG
mkdir -p "$OUT/obj" "$OUT/output/cal"
JSON=$REPO/oracle/cases/$CASE.json
if [ "$OUTJ" = 1 ]; then JSON=$REPO/oracle/cases/${CASE}_out.json; fi   # field output on (OutDAT, OutVTI) with several output formats
if [ "$PP" = 1 ]; then JSON=$REPO/oracle/cases/${CASE}_pp.json; [ -f "$JSON" ] || { echo "no $JSON"; exit 2; }; fi

DEFS=(-D__ACPP__ -DUSE_CXX_BOOST=1 -DUSE_DOUBLE -DSCHEME_ORDER=$WENO -DEIGEN_ALLOC=0 -D__SYNC_TIMER_=1
      -DESTIM_NAN=1 -DESTIM_OUT=0 -DThermo=1 -DArtificial_type=$AT
      "-DSelectDv=\"host\"" "-DINI_SAMPLE=\"$SAMPLE\"" "-DRFile=\"/runtime.dat/$MIX\"" "-DRPath=\"/runtime.dat\""
      "-DIniFile=\"$JSON\"")
if [ "$VISC" = 1 ]; then DEFS+=(-DVisc=1 -DVisc_Heat=1 -DVisc_Diffu=1); fi
if [ $COP = 1 ]; then DEFS+=(-DCOP -DPOSP=0 -DCOP_CHEME=0)
else DEFS+=(-DPOSP=0 -DNUM_REA=1 -DNUM_COP=0 -DCOP_CHEME=0 -DNUM_SPECIES=1 -DNCOP_Gamma=1.4); fi

INCS=(-I"$HERE/shim" -I"$REF/runtime.dat/$MIX" -I"$REF/$SDIR"
      -I"$REF/src/solver_Reconstruction/viscosity/Fourth_Order"
      -I"$REF/src/solver_Reconstruction/FDM_Method" -I"$REF/src/solver_Reconstruction/FDM_Method/positive-definite_eigen"
      -I"$REF/external" -I"$REF/src" -I"$REF/src/include"
      -I"$REF/src/solver_Reconstruction" -I"$REF/src/solver_Reconstruction/schemes" -I"$REF/src/solver_Reconstruction/viscosity"
      -I"$REF/src/solver_GetDt" -I"$REF/src/solver_Reaction")

if [ "$MODE" = parity ]; then OPT=(-O2 -ffp-contract=off)
else OPT=(-O3 -march=x86-64-v3 -fopenmp); fi   # x86-64-v3 (AVX2+FMA): the binary travels to the GPU box, whose CPU may differ
CXXFLAGS=(-std=c++17 -fpermissive -w -U_FORTIFY_SOURCE -D_FORTIFY_SOURCE=0 "${OPT[@]}")

SRCS=("$REF"/src/Fluids.cpp "$REF"/src/XFLUIDS.cpp "$REF"/src/read_ini/src/*.cpp
      "$REF"/src/read_ini/settings/read_json.cpp "$REF"/src/read_ini/outformat/outformat.cpp
      "$REF"/src/read_ini/inishape/inishape.cpp "$REF"/src/read_grid/readgrid.cpp
      "$REF"/src/solver_Ini/Ini_block.cpp "$REF"/src/solver_BCs/BCs_block.cpp
      "$REF"/src/solver_UpdateStates/UpdateStates_block.cpp
      "$REF"/external/timer/timer.cpp "$REF"/external/strsplit/strsplit.cpp "$REF"/external/ndassign/ndassign.cpp
      "$HERE"/ref_driver.cpp)

pids=(); objs=()
for s in "${SRCS[@]}"; do
  o="$OUT/obj/$(basename "${s%.cpp}").o"; objs+=("$o")
  g++ "${CXXFLAGS[@]}" "${DEFS[@]}" "${INCS[@]}" -c "$s" -o "$o" &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
g++ "${OPT[@]}" "${objs[@]}" -o "$OUT/XFLUIDS"
rm -rf "$OUT/obj"
echo "built $OUT/XFLUIDS"
