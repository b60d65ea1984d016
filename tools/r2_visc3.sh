#!/usr/bin/env bash
set -u
O=gpurun_out/r02d; mkdir -p $O
timeout 900 python -m pytest tests/test_xf_log.py tests/test_gpu_visc.py -x -q -m gpu > $O/test_visc.log 2>&1; echo "exp/pow + visc tests rc=$?"; tail -15 $O/test_visc.log
python bench.py --steps 5 --warmup 3 --visc 1 --no-cpu --e2e-steps 0 > $O/r02_bench_w5_visc_n1.json 2> $O/bench_w5visc.err
for f in $O/r02_bench_*.json; do python -c "
import json,sys
r=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', r['value'], r['ms_per_step'], r['roofline']['step_breakdown_ms'])"; done
