#!/usr/bin/env bash
set -u
O=gpurun_out/r02c; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_visc.py tests/test_gpu_output.py -x -q -m gpu > $O/test_visc.log 2>&1; echo "visc tests rc=$?"; tail -3 $O/test_visc.log
python bench.py --steps 5 --warmup 3 --visc 1 --no-cpu --e2e-steps 0 > $O/r02_bench_w5_visc_n1.json 2> $O/bench_w5visc.err
python bench.py --steps 5 --warmup 3 --weno 6 --pp 1 --alpha GLF --visc 1 --no-cpu --e2e-steps 0 > $O/r02_bench_preset_visc_n1.json 2> $O/bench_preset.err
ncu --set full --clock-control none -k 'regex:k_visc_flux3|k_transport|k_vde$|k_yi_minmax' -s 4 -c 4 -o $O/r02_full_visc -f \
    python bench.py --steps 1 --warmup 3 --visc 1 --no-cpu --e2e-steps 0 --profile-steps 0 > $O/ncu_visc.log 2>&1
python tools/ncu_summary.py full $O/r02_full_visc.ncu-rep > $O/r02_ncu_full_visc.md 2>/dev/null
rm -f $O/*.ncu-rep
for f in $O/r02_bench_*.json; do python -c "
import json,sys
r=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', r['value'], r['ms_per_step'], r['roofline']['step_breakdown_ms'])"; done
grep -E "^## |duration|dram read|dram write|FP64 pipe|top stall" $O/r02_ncu_full_visc.md | cut -c1-170
