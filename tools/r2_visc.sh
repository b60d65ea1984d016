#!/usr/bin/env bash
# viscous block after the k_transport rewrite: parity, bench lines, one ncu capture; plus the executed FP64 instruction counts of the sweeps of every benched workload
set -u
O=gpurun_out/r02b; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_visc.py tests/test_gpu_parity.py -x -q -m gpu -k "visc or fusion" > $O/test_visc.log 2>&1; echo "visc+fusion tests rc=$?"; tail -3 $O/test_visc.log
python bench.py --steps 5 --warmup 3 --visc 1 --no-cpu --e2e-steps 0 > $O/r02_bench_w5_visc_n1.json 2> $O/bench_w5visc.err
python bench.py --steps 5 --warmup 3 --weno 6 --pp 1 --alpha GLF --visc 1 --no-cpu --e2e-steps 0 > $O/r02_bench_preset_visc_n1.json 2> $O/bench_preset.err
ncu --set full --clock-control none -k 'regex:k_visc_flux|k_transport|k_vde' -s 8 -c 6 -o $O/r02_full_visc -f \
    python bench.py --steps 1 --warmup 3 --visc 1 --no-cpu --e2e-steps 0 --profile-steps 0 > $O/ncu_visc.log 2>&1
python tools/ncu_summary.py full $O/r02_full_visc.ncu-rep > $O/r02_ncu_full_visc.md 2>/dev/null
M=smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,gpu__time_duration.sum
run_ops() { # tag, bench args...
  local tag=$1; shift
  ncu --metrics $M --clock-control none -k 'regex:k_sweep' -s 6 -c 3 --csv --log-file $O/r02_fp64ops_$tag.csv \
      python bench.py --steps 1 --warmup 3 --no-cpu --e2e-steps 0 --profile-steps 0 "$@" > $O/ncu_ops_$tag.log 2>&1
}
run_ops sbi512
run_ops w7 --weno 7
run_ops cu6pp --weno 6 --pp 1
run_ops jet --workload jet
run_ops riemann --workload riemann
run_ops vortex --workload vortex
rm -f $O/*.ncu-rep $O/*.err
du -sh gpurun_out
for f in $O/r02_bench_*.json; do python -c "
import json,sys
r=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', r['value'], r['ms_per_step'], r['roofline']['step_breakdown_ms'])"; done
grep -E "^## |duration" $O/r02_ncu_full_visc.md | paste - - | cut -c1-160
