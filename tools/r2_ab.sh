#!/usr/bin/env bash
# A/B pass on the GPU box: parity suite, then the default bench with the marching (default) and the tiled round-1 path (XF_TILED=1).
# usage (under gpurun, from the repo root): bash tools/r2_ab.sh <tag> [pytest-args]
set -u
TAG=${1:-r2}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x ${2:-} 2>&1 | tail -30 > gpurun_out/${TAG}_pytest_gpu.log
tail -6 gpurun_out/${TAG}_pytest_gpu.log
python bench.py --steps 5 --warmup 3 --no-cpu --e2e-steps 2 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_bench.err
XF_MARCH=1 python bench.py --steps 5 --warmup 3 --no-cpu --e2e-steps 0 > gpurun_out/${TAG}_bench_march.json 2> gpurun_out/${TAG}_bench_march.err
python - <<PY
import json
for f in ("gpurun_out/${TAG}_bench.json", "gpurun_out/${TAG}_bench_march.json"):
    try:
        d = json.load(open(f))
        print(f, round(d["value"], 1), d["e2e"] and round(d["e2e"]["value"], 1), {k: round(v, 2) for k, v in d["roofline"]["step_breakdown_ms"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
