#!/usr/bin/env bash
# bench every tuning build under xfluids_b200/variants (default workload unless args are given); one summary line each
mkdir -p gpurun_out
for so in xfluids_b200/variants/lib_*.so; do
  n=$(basename $so .so)
  XF_LIB=$PWD/$so python bench.py --steps 3 --warmup 3 --no-cpu --e2e-steps 0 --profile-steps 1 "$@" > gpurun_out/var_$n.json 2> gpurun_out/var_$n.err
  python - "$n" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.load(open("gpurun_out/var_%s.json" % n))
    print(n, round(d["value"], 1), {k: round(v, 1) for k, v in d["roofline"]["step_breakdown_ms"].items() if k in ("prim", "sweep_x", "sweep_y", "sweep_z", "lu_rk", "step")})
except Exception as e:
    print(n, "ERR", e, open("gpurun_out/var_%s.err" % n).read()[-300:])
PY
done
