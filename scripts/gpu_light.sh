#!/bin/bash
# bench (both arms, both coordinate variants) + ncu launch list + one full capture; no pytest.  Usage: bash scripts/gpu_light.sh <tag>
TAG=${1:-light}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
timeout 120 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "ref rc=$?"
timeout 120 python bench.py --coords normalised --no-cpu-baseline > $OUT/bench_normalised.json 2> $OUT/bench_normalised.err; echo "norm rc=$?"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1; echo "ncu list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:power_ -s 2 -c 2 -f -o $OUT/prof \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?"
