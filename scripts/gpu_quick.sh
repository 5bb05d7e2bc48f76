#!/bin/bash
# quick GPU iteration: selected tests + bench variants.  Usage: bash scripts/gpu_quick.sh <tag> [pytest -k expr]
TAG=${1:-q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
if [ -n "$2" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q -k "$2" > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_gpu.log
fi
for c in raw normalised; do
  timeout 300 python bench.py --coords $c --no-cpu-baseline --steps 10 > $OUT/bench_$c.json 2> $OUT/bench_$c.err; echo "bench $c rc=$?"
  python - <<PY
import json
l=json.load(open("$OUT/bench_$c.json"))
k=l["roofline"]["kernels"]
print("$c", "step %.3f ms  fwd %.3f  bwd %.3f  e2e %.3f ms" % (l["ms_per_step"], k["power_fwd_kernel"]["ms"], k["power_bwd_kernel"]["ms"], l["e2e"]["ms_per_step"]))
PY
done
timeout 600 python scripts/bench_configs.py > $OUT/bench_configs.jsonl 2> $OUT/bench_configs.err; echo "configs rc=$?"; cut -c1-220 $OUT/bench_configs.jsonl
