#!/bin/bash
# fast GPU iteration: a subset of the parity tests + the bench on both coordinate variants.
# Usage: bash scripts/gpu_iter.sh <tag> [pytest -k expr]
TAG=${1:-it}
OUT=gpurun_out/$TAG
mkdir -p $OUT
K=${2:-"cull or hard_masks or vjp_against or slices or point_to_point or smooth_validity"}
timeout 900 python -m pytest tests -m gpu -x -q -k "$K" > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu.log
for c in raw normalised; do
  timeout 300 python bench.py --coords $c --no-cpu-baseline --steps 10 > $OUT/bench_$c.json 2> $OUT/bench_$c.err; echo "bench $c rc=$?"
  python - <<PY
import json
l=json.load(open("$OUT/bench_$c.json"))
k=l["roofline"]["kernels"]
print("$c", "step %.3f ms  fwd %.3f  bwd %.3f  e2e %.3f ms" % (l["ms_per_step"], k["power_fwd_kernel"]["ms"], k["power_bwd_kernel"]["ms"], l["e2e"]["ms_per_step"]))
PY
done
