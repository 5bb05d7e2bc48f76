#!/bin/bash
# bench on both coordinate variants, the other BASELINE configs, and one ncu capture of the Fermat / MinPath forward kernels.
# Usage: bash scripts/gpu_cfgs.sh <tag>
TAG=${1:-cfgs}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for c in raw normalised; do
  timeout 300 python bench.py --coords $c --no-cpu-baseline --steps 50 > $OUT/bench_$c.json 2> $OUT/bench_$c.err; echo "bench $c rc=$?"
  python - <<PY
import json
l=json.load(open("$OUT/bench_$c.json"))
k=l["roofline"]["kernels"]
print("$c", "step %.3f ms  fwd %.3f  bwd %.3f  e2e %.3f ms" % (l["ms_per_step"], k["power_fwd_kernel"]["ms"], k["power_bwd_kernel"]["ms"], l["e2e"]["ms_per_step"]))
PY
done
timeout 600 python scripts/bench_configs.py > $OUT/bench_configs.jsonl 2> $OUT/bench_configs.err; echo "configs rc=$?"; cut -c1-200 $OUT/bench_configs.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:power_fwd -s 3 -c 2 -f -o $OUT/cfg4 \
    python scripts/bench_configs.py --only cfg4 --steps 1 > $OUT/ncu_cfg4.log 2>&1; echo "ncu cfg4 rc=$?"
ls -la $OUT
