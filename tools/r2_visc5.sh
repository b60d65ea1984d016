#!/usr/bin/env bash
set -u
O=gpurun_out/r02f; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_visc.py -q -m gpu > $O/test_visc.log 2>&1; echo "visc tests rc=$?"; tail -3 $O/test_visc.log
python bench.py --steps 5 --warmup 3 --visc 1 --no-cpu --e2e-steps 0 > $O/r02_bench_w5_visc_n1.json 2> $O/bench_w5visc.err
python -c "
import json,sys
r=json.loads(open('$O/r02_bench_w5_visc_n1.json').read().strip().splitlines()[-1]); print('visc', r['value'], r['ms_per_step'], r['roofline']['step_breakdown_ms'])"
ncu --set full --clock-control none -k 'regex:k_transport' -s 2 -c 1 -o $O/r02_full_transport -f \
    python bench.py --steps 1 --warmup 3 --visc 1 --no-cpu --e2e-steps 0 --profile-steps 0 > $O/ncu_visc.log 2>&1
python tools/ncu_summary.py full $O/r02_full_transport.ncu-rep > $O/r02_ncu_transport.md 2>/dev/null
rm -f $O/*.ncu-rep
grep -E "^## |duration|regs/thread|achieved occ|issue slots|FP64 pipe|top stall" $O/r02_ncu_transport.md | cut -c1-200
