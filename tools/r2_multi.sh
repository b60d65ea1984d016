#!/usr/bin/env bash
# N-GPU pass: the whole GPU suite (the 2-GPU slab tests run for real here), the N = 1 bench, then the bench under torchrun on N GPUs
set -u
N=${1:-2}; TAG=${2:-r2m}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/${TAG}_pytest_gpu.log
tail -5 gpurun_out/${TAG}_pytest_gpu.log
python bench.py --steps 5 --warmup 3 --no-cpu --e2e-steps 2 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
tail -2 gpurun_out/${TAG}_bench_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
tail -5 gpurun_out/${TAG}_bench_n$N.err
python - <<PY
import json
for f in ("gpurun_out/${TAG}_bench_n1.json", "gpurun_out/${TAG}_bench_n$N.json"):
    try:
        d = json.load(open(f))
        print(f, round(d["value"], 1), "e2e", d["e2e"] and {k: (round(v, 1) if isinstance(v, float) else v) for k, v in d["e2e"].items() if k not in ("api", "limiter", "unit")},
              "check", d.get("slab_check"), "strong", d.get("strong") and {k: v for k, v in d["strong"].items() if k in ("value", "n1_value", "efficiency_vs_n1")})
        if "roofline" in d: print("   ", {k: round(v, 2) for k, v in d["roofline"]["step_breakdown_ms"].items()}, "frac_issue", d["roofline"].get("frac_issue"))
    except Exception as e:
        print(f, "ERR", e)
PY
