#!/usr/bin/env bash
# the driver's scaling run at N GPUs: reference arm first, then the bench under torchrun
set -u
N=${1:-8}; TAG=${2:-r2n8}
mkdir -p gpurun_out
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 5 --warmup 3 ) > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
tail -4 gpurun_out/${TAG}_bench_n$N.err
python - <<PY
import json
f = "gpurun_out/${TAG}_bench_n$N.json"
try:
    lines = [l for l in open(f).read().splitlines() if l.strip()]
    print("stdout lines:", len(lines))
    d = json.loads(lines[-1])
    print(round(d["value"], 1), "e2e", d["e2e"] and {k: (round(v, 1) if isinstance(v, float) else v) for k, v in d["e2e"].items() if k not in ("api", "limiter", "unit")},
          "check", d.get("slab_check"), "strong", d.get("strong") and {k: v for k, v in d["strong"].items() if k in ("value", "n1_value", "efficiency_vs_n1")}, "launches", d["gpu_launches"])
except Exception as e:
    print(f, "ERR", e)
PY
