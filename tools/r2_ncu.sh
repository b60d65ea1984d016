#!/usr/bin/env bash
# one `ncu --set full` capture of selected kernels of the default bench (SBI 512^3 unless XF_GRID is set)
# usage: bash tools/r2_ncu.sh <tag> <kernel-regex> [skip] [count] [extra bench args]
set -u
TAG=$1; RX=$2; SKIP=${3:-6}; CNT=${4:-3}; shift 4 || true
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k "regex:$RX" -s $SKIP -c $CNT -o gpurun_out/${TAG} -f \
    python bench.py --steps 1 --warmup 3 --no-cpu --e2e-steps 0 --profile-steps 0 "$@" > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log
python tools/ncu_summary.py full gpurun_out/${TAG}.ncu-rep > gpurun_out/${TAG}_summary.md
ncu -i gpurun_out/${TAG}.ncu-rep --page source --csv > gpurun_out/${TAG}_source.csv 2>/dev/null
grep -E "^## |duration|occ. limit|achieved occ|issue slots|FP64 pipe|local loads|dram read|dram write|top stall|L2 hit" gpurun_out/${TAG}_summary.md
