#!/usr/bin/env bash
# One GPU-box pass that produces every round-2 file under profiles/ (run under gpurun from the repo root; results land in gpurun_out/r02/).
set -u
O=gpurun_out/r02; mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu > $O/r02_pytest_gpu.log 2>&1; echo "gpu suite rc=$?" | tee $O/r02_pytest_gpu.rc; tail -3 $O/r02_pytest_gpu.log
python tools/peaks.py $O/r02_peaks.json > /dev/null 2> $O/peaks.err
python bench.py --steps 10 --warmup 3 > $O/r02_bench_n1.json 2> $O/bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_ref.json 2> $O/bench_ref.err
python bench.py --steps 5 --warmup 3 --fp 1 --no-cpu --e2e-steps 0 > $O/r02_bench_fast_n1.json 2> $O/bench_fast.err
python bench.py --steps 5 --warmup 3 --weno 7 --no-cpu --e2e-steps 0 > $O/r02_bench_w7_n1.json 2> $O/bench_w7.err
python bench.py --steps 5 --warmup 3 --weno 6 --pp 1 --no-cpu --e2e-steps 0 > $O/r02_bench_cu6pp_n1.json 2> $O/bench_cu6pp.err
python bench.py --steps 5 --warmup 3 --weno 6 --pp 1 --alpha GLF --visc 1 --no-cpu --e2e-steps 0 > $O/r02_bench_preset_visc_n1.json 2> $O/bench_preset.err
python bench.py --steps 5 --warmup 3 --visc 1 --no-cpu --e2e-steps 0 > $O/r02_bench_w5_visc_n1.json 2> $O/bench_w5visc.err
python bench.py --workload jet --steps 5 --warmup 3 --no-cpu --e2e-steps 0 > $O/r02_bench_jet_n1.json 2> $O/bench_jet.err
python bench.py --workload riemann --steps 20 --warmup 5 --no-cpu --e2e-steps 0 > $O/r02_bench_riemann_n1.json 2> $O/bench_riemann.err
python bench.py --workload vortex --steps 100 --warmup 10 --no-cpu --e2e-steps 0 > $O/r02_bench_vortex_n1.json 2> $O/bench_vortex.err
XF_MARCH=1 python bench.py --steps 5 --warmup 3 --no-cpu --e2e-steps 0 > $O/r02_bench_march_n1.json 2> $O/bench_march.err
# launch list of the default command (cold-cache, serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_sbi512.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 0 --profile-steps 0 > $O/ncu_list.log 2>&1
# full captures: the default hot kernels; WENO7 sweep; the 2-D configs
ncu --set full --clock-control none --import-source on -k 'regex:k_sweep|k_prim|k_rk' -s 12 -c 6 -o $O/r02_full_sbi512 -f \
    python bench.py --steps 1 --warmup 3 --no-cpu --e2e-steps 0 --profile-steps 0 > $O/ncu_full.log 2>&1
ncu --set full --clock-control none -k 'regex:k_sweep' -s 9 -c 3 -o $O/r02_full_w7 -f \
    python bench.py --steps 1 --warmup 3 --weno 7 --no-cpu --e2e-steps 0 --profile-steps 0 > $O/ncu_w7.log 2>&1
ncu --set full --clock-control none -k 'regex:k_sweep|k_prim|k_rk' -s 20 -c 4 -o $O/r02_full_riemann -f \
    python bench.py --workload riemann --steps 1 --warmup 3 --no-cpu --e2e-steps 0 --profile-steps 0 > $O/ncu_riemann.log 2>&1
ncu --set full --clock-control none -k 'regex:k_sweep|k_prim|k_rk' -s 20 -c 4 -o $O/r02_full_vortex -f \
    python bench.py --workload vortex --steps 1 --warmup 3 --no-cpu --e2e-steps 0 --profile-steps 0 > $O/ncu_vortex.log 2>&1
ncu --set full --clock-control none -k 'regex:k_transport|k_vde$|k_yi_minmax|k_sweep' -s 8 -c 6 -o $O/r02_full_visc -f \
    python bench.py --steps 1 --warmup 3 --visc 1 --no-cpu --e2e-steps 0 --profile-steps 0 > $O/ncu_visc.log 2>&1
# executed FP64 instruction counts of the sweeps of every benched workload (tools/make_traffic.py -> profiles/traffic.json -> roofline.frac_issue)
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
run_ops visc --visc 1
for r in sbi512 w7 riemann vortex visc; do python tools/ncu_summary.py full $O/r02_full_$r.ncu-rep > $O/r02_ncu_full_$r.md 2>/dev/null; done
ncu -i $O/r02_full_sbi512.ncu-rep --page source --csv > $O/r02_full_sbi512_source.csv 2>/dev/null
python tools/ncu_summary.py list $O/r02_launches_sbi512.csv > $O/r02_launches_sbi512.md 2>/dev/null
# gpurun brings back at most 64 MiB: keep the summaries, drop the raw reports, compress the source page
gzip -9 -f $O/r02_full_sbi512_source.csv
rm -f $O/*.ncu-rep.tmp $O/*.ncu-rep
du -sh gpurun_out
ls -la $O | head -60
