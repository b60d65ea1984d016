#!/usr/bin/env bash
set -u
O=gpurun_out/r02e; mkdir -p $O
timeout 900 python -m pytest tests/test_xf_log.py tests/test_gpu_visc.py tests/test_gpu_output.py -q -m gpu > $O/test_visc.log 2>&1; echo "tail on: rc=$?"; tail -4 $O/test_visc.log
XF_VISC_TAIL=0 timeout 900 python -m pytest tests/test_gpu_visc.py -q -m gpu > $O/test_visc_notail.log 2>&1; echo "tail off: rc=$?"; tail -3 $O/test_visc_notail.log
for t in 1 0; do
XF_VISC_TAIL=$t python bench.py --steps 5 --warmup 3 --visc 1 --no-cpu --e2e-steps 0 > $O/bench_w5_visc_tail$t.json 2> $O/bench_w5visc$t.err
python -c "
import json,sys
r=json.loads(open('$O/bench_w5_visc_tail$t.json').read().strip().splitlines()[-1]); print('tail=$t', r['value'], r['ms_per_step'], r['roofline']['step_breakdown_ms'])"
done
python bench.py --steps 5 --warmup 3 --weno 6 --pp 1 --alpha GLF --visc 1 --no-cpu --e2e-steps 0 > $O/r02_bench_preset_visc_n1.json 2> $O/bench_preset.err
python -c "
import json,sys
r=json.loads(open('$O/r02_bench_preset_visc_n1.json').read().strip().splitlines()[-1]); print('preset', r['value'], r['ms_per_step'], r['roofline']['step_breakdown_ms'])"
