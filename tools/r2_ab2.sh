#!/usr/bin/env bash
# parity suite + default bench (tiled sweeps with the TMA-staged conserved pencil) + the same with per-thread loads (variant notma)
set -u
TAG=${1:-r2}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/${TAG}_pytest_gpu.log
tail -4 gpurun_out/${TAG}_pytest_gpu.log
python bench.py --steps 5 --warmup 3 --no-cpu --e2e-steps 2 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_bench.err
bash tools/run_variants.sh
python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench.json"))
print("default", round(d["value"], 1), d["e2e"] and round(d["e2e"]["value"], 1), {k: round(v, 2) for k, v in d["roofline"]["step_breakdown_ms"].items()})
PY
