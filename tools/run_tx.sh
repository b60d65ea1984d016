set -u
mkdir -p gpurun_out; rm -f gpurun_out/tx.jsonl
run() { XF_LIB=$PWD/xfluids_b200/_variants/$2.so timeout 300 python bench.py --grid 512,256,256 --steps 3 --warmup 3 --no-cpu --e2e-steps 0 --profile-steps 2 --weno $3 --pp $4 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); b=d['roofline']['step_breakdown_ms']; print('$1', round(d['value'],1), round(b['sweep_x'],2), round(b['sweep_y'],2), round(b['sweep_z'],2))" | tee -a gpurun_out/tx.jsonl; }
for w in "5 1" "6 0" "6 1" "7 0" "7 1"; do set -- $w; run "txs_w$1_pp$2" txs $1 $2; run "base_w$1_pp$2" base $1 $2; done
