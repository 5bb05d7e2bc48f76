#!/bin/bash
# round-end evidence in one call: everything gpu_check.sh does + the other BASELINE configs + the normalised variant.
# Usage: bash scripts/gpu_final.sh <tag>
TAG=${1:-final}
bash scripts/gpu_check.sh $TAG
OUT=gpurun_out/$TAG
timeout 300 python bench.py --coords normalised --no-cpu-baseline > $OUT/bench_normalised.json 2> $OUT/bench_normalised.err; echo "bench normalised rc=$?"
timeout 600 python scripts/bench_configs.py > $OUT/bench_configs.jsonl 2> $OUT/bench_configs.err; echo "configs rc=$?"; cut -c1-200 $OUT/bench_configs.jsonl
