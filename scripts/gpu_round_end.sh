#!/bin/bash
# Round-end evidence in one gpurun call: full GPU parity suite, smoke, both bench arms, ncu launch list, ncu --set full
# captures of the headline step and of the dense leg, the other BASELINE configs, the normalised variant.
# Usage: bash scripts/gpu_round_end.sh <tag>
TAG=${1:-final}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
python -c "import jax" > $OUT/jax_probe.txt 2>&1 || true
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -4 $OUT/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log; tail -1 $OUT/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:power_ -s 6 -c 2 -f -o $OUT/prof_city \
    python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > $OUT/ncu_city.log 2>&1; echo "ncu city rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:power_ -s 6 -c 2 -f -o $OUT/prof_dense \
    python bench.py --only-dense --steps 1 > $OUT/ncu_dense.log 2>&1; echo "ncu dense rc=$?"
timeout 300 python bench.py --coords normalised --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_normalised.json 2> $OUT/bench_normalised.err; echo "bench normalised rc=$?"
timeout 600 python scripts/bench_configs.py > $OUT/bench_configs.jsonl 2> $OUT/bench_configs.err; echo "configs rc=$?"
python - <<PY
import json
l=json.load(open("$OUT/bench.json"))
print("step %.3f ms e2e %.3f ms frac %.3f dense %.3f+%.3f ms spot %s" % (l["ms_per_step"], l["e2e"]["ms_per_step"], l["roofline"]["frac"], l["roofline"]["dense_leg"]["fwd_ms"], l["roofline"]["dense_leg"]["bwd_ms"], json.dumps(l.get("parity_spotcheck",{}))[:300]))
PY
