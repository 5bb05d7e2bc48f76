#!/bin/bash
# ncu launch list + full capture of the power kernels on the bench workload.  Usage: bash scripts/gpu_profile.sh <tag> [raw|normalised]
TAG=${1:-prof}
COORDS=${2:-raw}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --coords $COORDS --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:power_ -s 2 -c 2 -f -o $OUT/prof \
    python bench.py --coords $COORDS --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la $OUT
