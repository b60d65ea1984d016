#!/usr/bin/env bash
# update -> recovery fusion: parity first, then A/B of the default bench with the fusion on and off (same box, back to back)
set -u
O=gpurun_out/fuse; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fusion" > $O/test_fusion.log 2>&1; echo "fusion tests rc=$?"; tail -5 $O/test_fusion.log
timeout 1500 python -m pytest tests -x -q -m gpu > $O/test_gpu.log 2>&1; echo "gpu suite rc=$?"; tail -4 $O/test_gpu.log
for f in 1 0; do
  XF_FUSE_PRIM=$f python bench.py --steps 8 --warmup 3 --no-cpu --e2e-steps 0 > $O/bench_fuse$f.json 2> $O/bench_fuse$f.err
  python - <<PY
import json
r=json.loads(open("$O/bench_fuse$f.json").read().strip().splitlines()[-1])
print("fuse=$f value", r["value"], "ms/step", r["ms_per_step"], "kernels", r.get("kernels_ms_per_step"), "launches", r.get("gpu_launches"))
PY
done
for w in jet riemann vortex; do for f in 1 0; do
  XF_FUSE_PRIM=$f python bench.py --workload $w --steps 10 --warmup 3 --no-cpu --e2e-steps 0 > $O/bench_${w}_fuse$f.json 2> $O/bench_${w}_fuse$f.err
  python -c "
import json
r=json.loads(open('$O/bench_${w}_fuse$f.json').read().strip().splitlines()[-1])
print('$w fuse=$f value', r['value'], 'ms/step', r['ms_per_step'], r.get('kernels_ms_per_step'))"
done; done
