#!/usr/bin/env bash
# last pass of the round on the final build: smoke, the whole GPU suite, the default bench line and the two viscous lines
set -u
O=gpurun_out/r02z; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
timeout 1500 python -m pytest tests -q -m gpu > $O/r02_pytest_gpu.log 2>&1; echo "gpu suite rc=$?"; tail -3 $O/r02_pytest_gpu.log
python bench.py > $O/r02_bench_default.json 2> $O/bench_default.err; tail -c 400 $O/r02_bench_default.json; echo
python bench.py --steps 5 --warmup 3 --weno 6 --pp 1 --alpha GLF --visc 1 --no-cpu --e2e-steps 0 > $O/r02_bench_preset_visc_n1.json 2> $O/bench_preset.err
python bench.py --steps 5 --warmup 3 --visc 1 --no-cpu --e2e-steps 0 > $O/r02_bench_w5_visc_n1.json 2> $O/bench_visc.err
for f in $O/r02_bench_*.json; do python -c "
import json,sys
r=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', r['value'], r['ms_per_step'], (r.get('e2e') or {}).get('value'), (r.get('roofline') or {}).get('frac_issue'))"; done
