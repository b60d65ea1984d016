#!/usr/bin/env bash
set -u
mkdir -p gpurun_out; rm -f gpurun_out/cu6.jsonl
run() { # name lib weno pp
  XF_LIB=$PWD/xfluids_b200/_variants/$2.so timeout 300 python bench.py --grid 512,256,256 --steps 3 --warmup 3 --no-cpu --e2e-steps 0 --profile-steps 2 --weno $3 --pp $4 2> gpurun_out/cu6_$1.err | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(json.dumps({'variant':'$1','value':d['value'],'ms_per_step':d['ms_per_step'],'breakdown':d['roofline']['step_breakdown_ms']}))" | tee -a gpurun_out/cu6.jsonl
}
run w5 base 5 0; run w5pp base 5 1; run w6 base 6 0; run w6pp base 6 1; run w7 base 7 0; run w7pp base 7 1
