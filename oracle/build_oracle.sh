#!/usr/bin/env bash
# oracle/build_oracle.sh -- TEST INFRASTRUCTURE: builds the CPU restatement into oracle/_build/liboracle.so
# (-ffp-contract=off so that no FMA is formed: same rounding sequence as the parity build of the reference)
set -euo pipefail
HERE=$(cd "$(dirname "$0")" && pwd)
mkdir -p "$HERE/_build"
g++ -std=c++17 -O2 -ffp-contract=off -fopenmp -fPIC -shared "$HERE/xf_oracle.cpp" -o "$HERE/_build/liboracle.so"
# throughput flavour for bench.py's cpu_baseline "port" leg (contraction allowed, like the reference's fast build)
g++ -std=c++17 -O3 -march=x86-64-v3 -fopenmp -fPIC -shared "$HERE/xf_oracle.cpp" -o "$HERE/_build/liboracle_fast.so"
# conditioning experiment: the same restatement with log() perturbed by 1 ulp on half of its arguments
g++ -std=c++17 -O2 -ffp-contract=off -fPIC -shared -DXO_PERTURB_LOG "$HERE/xf_oracle.cpp" -o "$HERE/_build/liboracle_plog.so"
echo "built $HERE/_build/liboracle.so"
