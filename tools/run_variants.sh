#!/usr/bin/env bash
# GPU-side tuning run: per-kernel step breakdown (bench.py's profile leg) for every library under xfluids_b200/_variants.
# usage: tools/run_variants.sh "<grid>" name1 name2 ...   -> gpurun_out/variants.jsonl
set -u
grid=$1; shift
mkdir -p gpurun_out
for n in "$@"; do
  XF_LIB=$PWD/xfluids_b200/_variants/$n.so timeout 300 python bench.py --grid "$grid" --steps 3 --warmup 3 --no-cpu --e2e-steps 0 --profile-steps 2 ${XF_BENCH_ARGS:-} 2> gpurun_out/var_$n.err | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(json.dumps({'variant':'$n','value':d['value'],'ms_per_step':d['ms_per_step'],'breakdown':d['roofline']['step_breakdown_ms']}))" | tee -a gpurun_out/variants.jsonl
done
