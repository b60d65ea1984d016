#!/usr/bin/env bash
# 2-D configs and the jet: marching (default) vs tiled (XF_TILED=1)
mkdir -p gpurun_out
for wl in riemann vortex jet; do
  for t in 0 1; do
    XF_MARCH=$((1-t)) python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu --e2e-steps 0 --profile-steps 1 > gpurun_out/r2d_${wl}_t$t.json 2> gpurun_out/r2d_${wl}_t$t.err
    python - $wl $t <<'PY'
import json, sys
wl, t = sys.argv[1:3]
try:
    d = json.load(open("gpurun_out/r2d_%s_t%s.json" % (wl, t)))
    print(wl, "tiled" if t == "1" else "march", round(d["value"], 1), {k: round(v, 2) for k, v in d["roofline"]["step_breakdown_ms"].items()})
except Exception as e:
    print(wl, t, "ERR", e, open("gpurun_out/r2d_%s_t%s.err" % (wl, t)).read()[-400:])
PY
  done
done
