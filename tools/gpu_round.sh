#!/usr/bin/env bash
# One GPU-box pass of the round's measurements (run under gpurun from the repo root): parity tests, the bench lines, the ncu launch
# list and one `ncu --set full` capture of the hot kernels.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py --steps 5 --warmup 3 --weno 6 --pp 1 --no-cpu --e2e-steps 0 > gpurun_out/bench_cu6pp_n1.json 2> gpurun_out/bench_cu6pp.err
python bench.py --steps 5 --warmup 3 --weno 7 --no-cpu --e2e-steps 0 > gpurun_out/bench_w7_n1.json 2> gpurun_out/bench_w7.err
python bench.py --workload jet --steps 5 --warmup 3 --no-cpu --e2e-steps 0 > gpurun_out/bench_jet_n1.json 2> gpurun_out/bench_jet.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 0 --profile-steps 0 > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:k_sweep|k_prim|k_rk' -s 12 -c 6 -o gpurun_out/prof_full -f \
    python bench.py --steps 1 --warmup 3 --no-cpu --e2e-steps 0 --profile-steps 0 > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/prof_full.ncu-rep --page raw --csv > gpurun_out/prof_full_raw.csv 2>/dev/null
ls -la gpurun_out/*.ncu-rep
